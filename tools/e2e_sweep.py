"""End-to-end (host buffers) fused scramble, variants next to the plain-copy ceilings of the same box, on 1..N
ranks at once (every rank its own GPU, all ranks copying concurrently, times = max over ranks):

    python tools/e2e_sweep.py                                      # one GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/e2e_sweep.py

Variants: host-supplied moves (30 B/instance in) or seeds (4 B/instance in, moves drawn on the device);
with or without the reward array on the way back (4 of 59 B/instance); cudaMallocHost buffers (torch
pin_memory) or 2 MiB huge-page buffers (ops.host_buffer)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from rubiks_cube_solver_b200 import dist as cdist
from rubiks_cube_solver_b200 import ops


def main():
    rank, local, world = cdist.init_from_env()
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    n, depth, S = 8 << 20, 30, 54
    iters = 5

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def timed(fn):
        for _ in range(2):
            fn()
        barrier()
        t0 = time.perf_counter()
        for _ in range(iters):
            fn()
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / iters
        dt = cdist.max_over_ranks(dt, dev)
        barrier()
        return dt

    def alloc(shape, dtype, huge):
        return ops.host_buffer(shape, dtype) if huge else torch.empty(shape, dtype=dtype).pin_memory()

    out = {"world": world, "n_per_gpu": n, "depth": depth}
    gen = torch.Generator().manual_seed(1234 + rank)
    for huge in (False, True):
        h_moves = alloc((n, depth), torch.uint8, huge)
        h_moves.copy_(torch.randint(0, 12, (n, depth), dtype=torch.uint8, generator=gen))
        h_seeds = alloc((n,), torch.int32, huge)
        h_seeds.copy_(torch.arange(n, dtype=torch.int32) + rank * n)
        h_states = alloc((n, S), torch.uint8, huge)
        h_solved = alloc((n,), torch.uint8, huge)
        h_reward = alloc((n,), torch.float32, huge)
        tag = "huge" if huge else "pinned"
        for chunk in (1 << 19, 1 << 20):
            pipe = ops.HostScramblePipeline(3, depth, chunk_instances=chunk, n_stages=3, device=dev)
            t = timed(lambda: pipe.run(h_moves, h_states, h_solved, h_reward))
            out["%s_moves_reward_chunk%d_ms" % (tag, chunk >> 10)] = t * 1e3
            t = timed(lambda: pipe.run(h_moves, h_states, h_solved, None, want_reward=False))
            out["%s_moves_noreward_chunk%d_ms" % (tag, chunk >> 10)] = t * 1e3
            t = timed(lambda: pipe.reset(h_seeds, h_states, h_solved, h_reward))
            out["%s_seeds_reward_chunk%d_ms" % (tag, chunk >> 10)] = t * 1e3
            t = timed(lambda: pipe.reset(h_seeds, h_states, h_solved, None, want_reward=False))
            out["%s_seeds_noreward_chunk%d_ms" % (tag, chunk >> 10)] = t * 1e3
            pipe.close()
        # plain-copy ceilings with the same buffers: D2H of the outputs alone, and with the H2D of the inputs beside it
        d_states = torch.empty((n, S), dtype=torch.uint8, device=dev)
        d_solved = torch.empty(n, dtype=torch.uint8, device=dev)
        d_reward = torch.empty(n, dtype=torch.float32, device=dev)
        d_moves = torch.empty((n, depth), dtype=torch.uint8, device=dev)
        d_seeds = torch.empty(n, dtype=torch.int32, device=dev)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def copies(h2d, reward):
            with torch.cuda.stream(s1):
                h_states.copy_(d_states, non_blocking=True)
                h_solved.copy_(d_solved, non_blocking=True)
                if reward:
                    h_reward.copy_(d_reward, non_blocking=True)
            with torch.cuda.stream(s2):
                if h2d == "moves":
                    d_moves.copy_(h_moves, non_blocking=True)
                elif h2d == "seeds":
                    d_seeds.copy_(h_seeds, non_blocking=True)
            s1.synchronize()
            s2.synchronize()

        for h2d in ("moves", "seeds", "none"):
            for reward in (True, False):
                t = timed(lambda: copies(h2d, reward))
                out["%s_ceiling_h2d_%s_%s_ms" % (tag, h2d, "reward" if reward else "noreward")] = t * 1e3
        del h_moves, h_seeds, h_states, h_solved, h_reward
    if rank == 0:
        print(json.dumps(out, indent=1), flush=True)
    if world > 1:
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
