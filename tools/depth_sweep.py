"""Fused-scramble throughput over depths (generic vs compile-time-depth instantiations of K1p, and the
single-move fallbacks beyond depth 96):  python tools/depth_sweep.py [sizes...]
AB_LIB=<another build of libcube_b200.so> sweeps that build; SWEEP_DEPTHS=64,128,320 restricts the depths."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rubiks_cube_solver_b200 import _lib

if os.environ.get("AB_LIB"):
    _lib.LIB_PATH = os.environ["AB_LIB"]
from rubiks_cube_solver_b200 import ops

dev = torch.device("cuda", 0)
only = [int(v) for v in sys.argv[1:]]                      # optional: the cube sizes to sweep
DEPTHS = (1, 2, 7, 8, 10, 16, 19, 20, 21, 24, 29, 30, 31, 32, 40, 48, 64, 96, 97, 128, 160, 200, 256, 320, 321, 480, 1000)
if os.environ.get("SWEEP_DEPTHS"):
    DEPTHS = tuple(int(v) for v in os.environ["SWEEP_DEPTHS"].split(","))
for size, a, n in ((3, 12, 4 << 20), (2, 6, 8 << 20)):
    if only and size not in only:
        continue
    s = ops.N_STICKERS[size]
    for depth in DEPTHS:
        if depth > 320:
            n = min(n, 1 << 20)
        moves = torch.randint(0, a, (n, depth), dtype=torch.uint8, device=dev)
        st = torch.empty((n, s), dtype=torch.uint8, device=dev)
        so = torch.empty(n, dtype=torch.uint8, device=dev)
        rw = torch.empty(n, dtype=torch.float32, device=dev)
        for _ in range(2):
            ops.scramble(size, moves, out=st, solved=so, reward=rw)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            ops.scramble(size, moves, out=st, solved=so, reward=rw)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        gbs = n * (depth + s + 5) / ms / 1e6
        print("size %d depth %3d: %8.4f ms  %.3e tr/s  %6.0f GB/s algorithmic (%.2f of 6553)" % (
            size, depth, ms, n * depth / ms * 1e3, gbs, gbs / 6553.3), flush=True)
        del moves
