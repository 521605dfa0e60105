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
for size, depths in ((3, (30, 31, 7, 45)), (2, (20, 9, 33))):
    a = ops.N_ACTIONS[size]
    for depth in depths:
        n = 64 * 40 + 17
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
torch.cuda.synchronize()
print("sanitize_small ok")
