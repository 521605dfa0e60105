"""Per-call latency of the drop-in CubeEnv (one cube per call, the way train.py / mcts.py / test.py drive it):
    python tools/env_latency.py [calls]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from rubiks_cube_solver_b200.env import make_env

calls = int(sys.argv[1]) if len(sys.argv) > 1 else 2000
for size in (2, 3):
    env = make_env("cuda", size)
    action_dim = env.action_dim
    env.reset(seed=1, scramble_count=10)
    rng = np.random.RandomState(0)
    acts = rng.randint(action_dim, size=calls)
    for a in acts[:50]:
        env.step(int(a))
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for a in acts:
        env.step(int(a))
    t1 = time.perf_counter()
    for i in range(200):
        env.reset(seed=i, scramble_count=20)
    t2 = time.perf_counter()
    from rubiks_cube_solver_b200 import ops
    hc = ops.host_cube(size, torch.cuda.current_device())
    st = np.ascontiguousarray(env.sim_cube, dtype=np.uint8)
    t3 = time.perf_counter()
    for a in acts:
        st, oh, done = hc.step(st, int(a))
    t4 = time.perf_counter()
    print("size %d: env.step %.1f us/call (HostCube.step alone %.1f), reset(seed, 20) %.1f us/call" % (
        size, (t1 - t0) / calls * 1e6, (t4 - t3) / calls * 1e6, (t2 - t1) / 200 * 1e6), flush=True)
