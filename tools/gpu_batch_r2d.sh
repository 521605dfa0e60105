#!/bin/bash
set -u
cd "$(dirname "$0")/.."
O=gpurun_out
for i in 1 2; do
CUBE_EARLY_WAIT=0 python tools/ab_k1p.py 2>&1 | tail -1
CUBE_EARLY_WAIT=1 python tools/ab_k1p.py 2>&1 | tail -1
done
timeout 600 python -m pytest tests -m gpu -x -q -k "scramble" > $O/r2_pytest_d.log 2>&1; echo "pytest rc=$?"; tail -2 $O/r2_pytest_d.log
python - <<'PY'
import sys, time, torch
sys.path.insert(0, '.')
from rubiks_cube_solver_b200 import ops
dev = torch.device('cuda', 0)
n, depth, S = 8 << 20, 30, 54
h_seeds = torch.arange(n, dtype=torch.int32).pin_memory()
h_states = torch.empty((n, S), dtype=torch.uint8).pin_memory()
h_solved = torch.empty(n, dtype=torch.uint8).pin_memory()
h_reward = torch.empty(n, dtype=torch.float32).pin_memory()
for chunk in (1 << 17, 1 << 18, 1 << 19, 1 << 20):
    for stages in (2, 3, 4):
        pipe = ops.HostScramblePipeline(3, depth, chunk_instances=chunk, n_stages=stages, device=dev)
        for _ in range(2):
            pipe.reset(h_seeds, h_states, h_solved, h_reward)
        t0 = time.perf_counter()
        for _ in range(5):
            pipe.reset(h_seeds, h_states, h_solved, h_reward)
        dt = (time.perf_counter() - t0) / 5
        print("seeds chunk %8d stages %d: %.3f ms" % (chunk, stages, dt * 1e3), flush=True)
        pipe.close()
PY
