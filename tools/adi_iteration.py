"""ADI iteration (cube_env.py:177-252) at BASELINE config 4's size, phases device-timed:
    python tools/adi_iteration.py [--cubes 139810] [--dtype bf16|f32] [--chunks 2048,4096,8192,16384,65536]
Prints one JSON object per chunk size: total / prefixes / expand / net / targets milliseconds, and the
stand-alone cube_expand time of the same number of parents beside them."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from rubiks_cube_solver_b200 import adi, ops


class DeepCubeNet(torch.nn.Module):                       # DeepCube's layer shapes (model.py:7-29), config.yaml hidden [1024, 256, 128]
    def __init__(self, d=480, a=12, hidden=(1024, 256, 128)):
        super().__init__()
        nn = torch.nn
        self.enc = nn.Sequential(nn.Flatten(), nn.Linear(d, hidden[0]), nn.ELU(), nn.Linear(hidden[0], hidden[1]), nn.ELU())
        self.pol = nn.Sequential(nn.Linear(hidden[1], hidden[2]), nn.ELU(), nn.Linear(hidden[2], a))
        self.val = nn.Sequential(nn.Linear(hidden[1], hidden[2]), nn.ELU(), nn.Linear(hidden[2], 1))

    def forward(self, x):
        if x.dim() == 2:
            x = x.unsqueeze(0)
        h = self.enc(x)
        return self.val(h), self.pol(h)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cubes", type=int, default=139810)
    ap.add_argument("--depth", type=int, default=30)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--chunks", default="2048,4096,8192,16384,65536")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    dtype = torch.bfloat16 if args.dtype == "bf16" else torch.float32
    torch.manual_seed(0)
    net = DeepCubeNet().to(dev).to(dtype)
    gen = torch.Generator(device=dev).manual_seed(5)
    moves = torch.randint(0, 12, (args.cubes, args.depth), dtype=torch.uint8, device=dev, generator=gen)
    p = args.cubes * args.depth
    # the expansion alone, whole batch resident (round 1's config-4 figure) when it fits
    parents = adi.scramble_prefixes(3, moves).view(p, 54)
    esize = 2 if dtype == torch.bfloat16 else 4
    expand_ms = None
    if p * 12 * 480 * esize < 120e9:
        child = torch.empty((p, 12, 20, 24), dtype=dtype, device=dev)
        for _ in range(2):
            ops.expand(3, parents, dtype=dtype, child_onehot=child)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            ops.expand(3, parents, dtype=dtype, child_onehot=child)
        e1.record()
        torch.cuda.synchronize()
        expand_ms = e0.elapsed_time(e1) / 3
        del child
        torch.cuda.empty_cache()
    for chunk in [int(c) for c in args.chunks.split(",")]:
        adi.generate_samples(3, moves[:2048], net, 1.0, forward_chunk=chunk)         # warm-up (cuBLAS heuristics)
        torch.cuda.synchronize()
        timers = {}
        t0 = time.perf_counter()
        out = adi.generate_samples(3, moves, net, 1.0, forward_chunk=chunk, timers=timers)
        torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        # untimed second run for the wall clock without the event bookkeeping
        t0 = time.perf_counter()
        adi.generate_samples(3, moves, net, 1.0, forward_chunk=chunk)
        torch.cuda.synchronize()
        wall2 = (time.perf_counter() - t0) * 1e3
        non_net = timers["prefixes_ms"] + timers["expand_ms"] + timers["targets_ms"]
        print(json.dumps({"parents": p, "dtype": args.dtype, "chunk": chunk, "wall_ms": wall2, "wall_ms_with_events": wall,
                          "phases": timers, "non_net_ms": non_net, "expand_alone_ms": expand_ms,
                          "non_net_over_expand": (non_net / expand_ms) if expand_ms else None,
                          "samples_per_s": p / wall2 * 1e3}), flush=True)
        del out


if __name__ == "__main__":
    main()
