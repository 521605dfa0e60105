"""End-to-end (host buffers) fused scramble over chunk sizes / stage counts, next to the raw PCIe copy
ceilings of the same box:  python tools/e2e_sweep.py"""
import sys, time
sys.path.insert(0, '/root/repo')
import torch
from rubiks_cube_solver_b200 import ops
dev = torch.device('cuda', 0)
n, depth, S = 8 << 20, 30, 54
moves = torch.randint(0, 12, (n, depth), dtype=torch.uint8, device=dev)
h_moves = moves.cpu().pin_memory()
h_states = torch.empty((n, S), dtype=torch.uint8).pin_memory()
h_solved = torch.empty(n, dtype=torch.uint8).pin_memory()
h_reward = torch.empty(n, dtype=torch.float32).pin_memory()
for chunk in (1 << 18, 1 << 19, 1 << 20, 1 << 21):
    for stages in (2, 3, 4):
        pipe = ops.HostScramblePipeline(3, depth, chunk_instances=chunk, n_stages=stages, device=dev)
        for _ in range(2):
            pipe.run(h_moves, h_states, h_solved, h_reward)
        t0 = time.perf_counter()
        for _ in range(5):
            pipe.run(h_moves, h_states, h_solved, h_reward)
        dt = (time.perf_counter() - t0) / 5
        print("chunk %8d stages %d: %.3f ms  %.3e tr/s  (%.1f GB/s D2H)" % (chunk, stages, dt * 1e3, n * depth / dt, n * 59 / dt / 1e9))
        pipe.close()
# raw copy ceilings
d = torch.empty(n * 59, dtype=torch.uint8, device=dev)
h = torch.empty(n * 59, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    h.copy_(d, non_blocking=True)
torch.cuda.synchronize()
print("plain D2H of 495 MB: %.3f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
d2 = torch.empty(n * 30, dtype=torch.uint8, device=dev)
h2 = torch.empty(n * 30, dtype=torch.uint8).pin_memory()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(s1):
        h.copy_(d, non_blocking=True)
    with torch.cuda.stream(s2):
        d2.copy_(h2, non_blocking=True)
torch.cuda.synchronize()
print("D2H 495 MB || H2D 252 MB: %.3f ms" % ((time.perf_counter() - t0) / 5 * 1e3))
