"""Small invocations of every kernel for compute-sanitizer runs:
    compute-sanitizer --tool memcheck|racecheck|synccheck python tools/sanitize_small.py"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from rubiks_cube_solver_b200 import ops
from oracle import cube_np as O

dev = torch.device("cuda", 0)
rng = np.random.RandomState(1)
ops.set_reserved_sms(ops.sm_count() - 2)           # two persistent CTAs: several tiles per warp, both move buffers
for size, depths in ((3, (30, 31, 7, 45, 32, 24, 64, 128)), (2, (20, 9, 33, 16, 32, 96))):
    a = ops.N_ACTIONS[size]
    for depth in depths:
        n = 64 * 200 + 17
        moves = rng.randint(a, size=(n, depth)).astype(np.uint8)
        st, so, rw = ops.scramble(size, torch.from_numpy(moves).to(dev))
        want = O.scramble(size, moves)
        assert (st.cpu().numpy() == want).all()
        assert (so.cpu().numpy().astype(bool) == O.is_solved(size, want)).all()
    act = rng.randint(a, size=n).astype(np.uint8)
    stepped, s2, _ = ops.step(size, st.clone(), torch.from_numpy(act).to(dev))
    want2 = O.apply_moves(size, want, act)
    assert (stepped.cpu().numpy() == want2).all()
    mv = rng.randint(a, size=(n, 5)).astype(np.uint8)
    walked, s3, _ = ops.walk(size, st, torch.from_numpy(mv).to(dev))
    assert (walked.cpu().numpy() == O.scramble(size, mv, init=want)).all()
    res = ops.expand(size, stepped[:300].contiguous(), dtype=torch.bfloat16, want_children=True, want_parent_onehot=True)
    wc, ws = O.expand(size, want2[:300])
    assert (res["children"].cpu().numpy() == wc).all()
    hc = ops.HostCube(size, max_depth=64)            # mapped pinned page: kernels read / write host memory
    hs, ho, _ = hc.scramble(moves[0, :min(depth, 64)].copy())
    assert (hs == O.scramble(size, moves[:1, :min(depth, 64)])[0]).all()
    hs2, _, _ = hc.step(hs, int(act[0]))
    assert (hs2 == O.apply_moves(size, hs[None], act[:1])[0]).all() and (hc.encode(hs) == ho).all()
    hc.close()
ops.set_reserved_sms(0)
torch.cuda.synchronize()
print("sanitize_small ok")
