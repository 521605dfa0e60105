"""Per-phase device time of the batched MCTS (bench.py's mcts_2x2_batched_search workload):
    python tools/mcts_phases.py [--trees 65536] [--sims 50]
traverse / expand_codes / net / softmax glue / update, CUDA events around every phase of every simulation."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

from rubiks_cube_solver_b200 import mcts_batch, ops


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--trees", type=int, default=65536)
    ap.add_argument("--sims", type=int, default=50)
    ap.add_argument("--depth", type=int, default=8)
    ap.add_argument("--variants", action="store_true", help="also time graph replay and a bf16 net")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    # the same net as bench.measure_mcts
    nn = torch.nn

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.enc = nn.Sequential(nn.Flatten(), nn.Linear(147, 512), nn.ELU(), nn.Linear(512, 128), nn.ELU())
            self.pol = nn.Sequential(nn.Linear(128, 64), nn.ELU(), nn.Linear(64, 6))
            self.val = nn.Sequential(nn.Linear(128, 64), nn.ELU(), nn.Linear(64, 1))

        def forward(self, x):
            if x.dim() == 2:
                x = x.unsqueeze(0)
            h = self.enc(x)
            return self.val(h), self.pol(h)

    torch.manual_seed(1)
    net = Net()
    fixture = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "pin222.npz")
    if os.path.exists(fixture):
        g = np.load(fixture)
        names = {"encoder_net": "enc", "policy_net": "pol", "value_net": "val"}
        net.load_state_dict({names[k[2:].split(".")[0]] + k[2 + len(k[2:].split(".")[0]):]: torch.from_numpy(g[k])
                             for k in g.files if k.startswith("w:")})
    net = net.to(dev)
    gen = torch.Generator(device=dev).manual_seed(77)
    roots, _, _ = ops.scramble(2, torch.randint(0, 6, (args.trees, args.depth), dtype=torch.uint8, device=dev, generator=gen),
                               want_flags=False)
    table = torch.randint(0, 6, (args.trees, 8 * (args.sims + 1)), generator=torch.Generator().manual_seed(3), dtype=torch.uint8)
    search = mcts_batch.BatchedMCTS(net, 2, num_sim=args.sims)
    search.run(roots[:256], rand_table=table[:256])
    torch.cuda.synchronize()
    for rep in range(2):
        timers = {}
        t0 = time.perf_counter()
        res = search.run(roots, rand_table=table, timers=timers if rep == 1 else None)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        print(json.dumps({"trees": args.trees, "sims": args.sims, "wall_ms": wall, "solved": int(res["solved"].sum()),
                          "sims_run": int(res["n_sims"].sum()), "phases_ms": timers}), flush=True)
    if args.variants:
        variants(net, roots, table, args.sims)


def variants(net, roots, table, sims):
    """Wall time of the whole search: eager / CUDA-graph replay, float32 / bfloat16 net."""
    import copy
    out = {}
    for name, dtype, graph in (("f32_eager", torch.float32, False), ("f32_graph", torch.float32, True),
                               ("bf16_eager", torch.bfloat16, False), ("bf16_graph", torch.bfloat16, True)):
        m = copy.deepcopy(net).to(dtype)
        search = mcts_batch.BatchedMCTS(m, 2, num_sim=sims, obs_dtype=dtype, graph=graph)
        search.run(roots[:256], rand_table=table[:256])
        torch.cuda.synchronize()
        best = None
        for _ in range(3):
            t0 = time.perf_counter()
            res = search.run(roots, rand_table=table)
            torch.cuda.synchronize()
            dt = (time.perf_counter() - t0) * 1e3
            best = dt if best is None else min(best, dt)
        out[name] = {"ms": best, "solved": int(res["solved"].sum()), "sims_run": int(res["n_sims"].sum())}
    print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
