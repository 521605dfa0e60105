"""Where a K1p launch's fixed cost goes: per-warp timestamps (%globaltimer) of an INSTRUMENTED build of the library
(AB_LIB=<path>; the kernel records entry / table filled / dependency wait over / first tile's moves landed / last
tile stored) for four back-to-back launches of config 3's slice (8 Mi x depth 30)."""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from rubiks_cube_solver_b200 import _lib

_lib.LIB_PATH = os.environ["AB_LIB"]
from rubiks_cube_solver_b200 import ops

dev = torch.device("cuda", 0)
n, d = int(os.environ.get("TL_N", 8 * 2 ** 20)), 30
moves = torch.randint(0, 12, (n, d), dtype=torch.uint8, device=dev, generator=torch.Generator(device=dev).manual_seed(1234))
st = torch.empty((n, 54), dtype=torch.uint8, device=dev)
so = torch.empty(n, dtype=torch.uint8, device=dev)
rw = torch.empty(n, dtype=torch.float32, device=dev)
for _ in range(12):
    ops.scramble(3, moves, out=st, solved=so, reward=rw)
torch.cuda.synchronize()
raw = ctypes.CDLL(_lib.LIB_PATH)
buf = np.zeros((4, 148, 32, 6), dtype=np.uint64)
assert raw.cube_debug_timeline(buf.ctypes.data_as(ctypes.c_void_p)) == 0
warps = int((buf[0, 0, :, 0] != 0).sum())
buf = buf[:, :, :warps, :].astype(np.int64)
order = np.argsort(buf[:, :, :, 0].min(axis=(1, 2)))
t0 = buf[order[0], :, :, 0].min()
print("warps per CTA", warps, " (times in us from the first launch's first CTA entry)")
names = ["entry", "filled", "wait over", "first tile", "done"]
prev_done = None
for k in order:
    b = (buf[k, :, :, :5] - t0) / 1e3
    print("launch", int(k))
    for j, nm in enumerate(names):
        v = b[:, :, j].ravel()
        print("   %-10s min %8.2f  p10 %8.2f  median %8.2f  p90 %8.2f  max %8.2f" % (nm, v.min(), np.percentile(v, 10), np.median(v), np.percentile(v, 90), v.max()))
    sm_done = b[:, :, 4].max(axis=1)                      # per CTA: its last warp
    sm_first = b[:, :, 4].min(axis=1)                     # per CTA: its first warp to run out of tiles
    print("   per CTA: last warp done - first warp done: median %.2f max %.2f us;  fill %.2f us;  entry -> first tile median %.2f us" % (
        np.median(sm_done - sm_first), (sm_done - sm_first).max(), np.median(b[:, :, 1] - b[:, :, 0]), np.median(b[:, :, 3] - b[:, :, 0])))
    # idle warp-time at the tail: sum over warps of (kernel end - warp done) / (warps * kernel span)
    end = b[:, :, 4].max()
    start = b[:, :, 0].min()
    print("   span %.2f us; warp-time idle before the grid's end: %.2f %% of span; between entry and first tile: %.2f %%" % (
        end - start, 100 * (end - b[:, :, 4]).mean() / (end - start), 100 * (b[:, :, 3] - start).mean() / (end - start)))
    if prev_done is not None:
        smid = buf[k, :, 0, 5]
        gaps = []
        for c in range(148):
            if smid[c] in prev_done:
                gaps.append(b[c, :, 0].min() - prev_done[smid[c]])
        print("   same SM: this CTA's entry - previous launch's CTA's last warp done: median %.2f  min %.2f  max %.2f us" % (np.median(gaps), min(gaps), max(gaps)))
    prev_done = {int(buf[k, c, 0, 5]): b[c, :, 4].max() for c in range(148)}
