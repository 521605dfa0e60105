"""Fixed overhead of a 5-step timed region with / without the NVML sampler thread and a second sync
(used to move the sampler thread's start-up out of bench.py's timed region)."""
import sys, time, threading
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch, bench
from rubiks_cube_solver_b200 import ops
dev = torch.device('cuda', 0)
n, depth = 8 << 20, 30
moves = torch.randint(0, 12, (n, depth), dtype=torch.uint8, device=dev)
st = torch.empty((n, 54), dtype=torch.uint8, device=dev)
so = torch.empty(n, dtype=torch.uint8, device=dev)
rw = torch.empty(n, dtype=torch.float32, device=dev)
cb = torch.zeros((64, 4), dtype=torch.int64, device=dev)
def run(K, sampler, sync_twice):
    for _ in range(3):
        ops.scramble(3, moves, out=st, solved=so, reward=rw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ctx = bench.ClockSampler(0) if sampler else None
    if ctx: ctx.__enter__()
    if sync_twice: torch.cuda.synchronize()
    e0.record()
    for k in range(K):
        ops.scramble(3, moves, out=st, solved=so, reward=rw, counters=cb[k])
    e1.record()
    torch.cuda.synchronize()
    if ctx: ctx.__exit__(None, None, None)
    return e0.elapsed_time(e1) / K
for sampler in (False, True):
    for sync_twice in (False, True):
        print("sampler", sampler, "sync_twice", sync_twice, ["%.4f" % run(5, sampler, sync_twice) for _ in range(4)])
