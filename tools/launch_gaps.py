"""Per-launch gaps of chained scramble launches (events between launches break the dependent-launch
chain on purpose): how much of a short timed region is launch latency."""
import sys, time
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch
from rubiks_cube_solver_b200 import ops
dev = torch.device('cuda', 0)
n, depth = 8 << 20, 30
moves = torch.randint(0, 12, (n, depth), dtype=torch.uint8, device=dev)
st = torch.empty((n, 54), dtype=torch.uint8, device=dev)
so = torch.empty(n, dtype=torch.uint8, device=dev)
rw = torch.empty(n, dtype=torch.float32, device=dev)
cb = torch.zeros((64, 4), dtype=torch.int64, device=dev)
for _ in range(5):
    ops.scramble(3, moves, out=st, solved=so, reward=rw)
torch.cuda.synchronize()
for trial in range(3):
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(7)]
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    ev[0].record()
    host = []
    for k in range(6):
        ops.scramble(3, moves, out=st, solved=so, reward=rw, counters=cb[k])
        ev[k + 1].record()
        host.append(time.perf_counter() - t0)
    torch.cuda.synchronize()
    print("gpu ms between events:", ["%.4f" % ev[k].elapsed_time(ev[k + 1]) for k in range(6)], "host us:", ["%.0f" % (h * 1e6) for h in host])
# without per-step event records
for K in (5, 20):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K):
        ops.scramble(3, moves, out=st, solved=so, reward=rw, counters=cb[k])
    e1.record()
    torch.cuda.synchronize()
    print(K, "steps:", e0.elapsed_time(e1) / K, "ms/step")
