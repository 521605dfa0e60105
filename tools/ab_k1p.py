"""A/B timing of the fused scramble (config 3, 8 Mi x depth 30) in back-to-back launches, as bench.py's timed loop
runs it.  AB_LIB=<path to another build of libcube_b200.so> times that build instead of the tree's (same box, same
process layout): e.g. a build of HEAD exported with `git archive` next to the working tree's."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from rubiks_cube_solver_b200 import _lib

if os.environ.get("AB_LIB"):
    _lib.LIB_PATH = os.environ["AB_LIB"]
from rubiks_cube_solver_b200 import ops

dev = torch.device("cuda", 0)
n, d = 8 * 2 ** 20, int(os.environ.get("AB_DEPTH", "30"))
moves = torch.randint(0, 12, (n, d), dtype=torch.uint8, device=dev, generator=torch.Generator(device=dev).manual_seed(1234))
st = torch.empty((n, 54), dtype=torch.uint8, device=dev)
so = torch.empty(n, dtype=torch.uint8, device=dev)
rw = torch.empty(n, dtype=torch.float32, device=dev)
best = []
for rep in range(5):
    for _ in range(5):
        ops.scramble(3, moves, out=st, solved=so, reward=rw)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(200):
        ops.scramble(3, moves, out=st, solved=so, reward=rw)
    e1.record()
    torch.cuda.synchronize()
    best.append(e0.elapsed_time(e1) / 200)
ms = sorted(best)[len(best) // 2]
print("lib=%s NBUF=%s depth %d  median %.5f ms  min %.5f  %.4e tr/s  frac %.4f" % (
    os.environ.get("AB_LIB", "tree"), os.environ.get("CUBE_PAIR_NBUF", "-"), d, ms, min(best), n * d / ms * 1e3,
    n * (d + 59) / ms / 1e6 / 6553.3))
