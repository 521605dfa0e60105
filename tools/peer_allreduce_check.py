"""Under torchrun (one rank per GPU of one box): the peer-memory all-reduce of the counters (cube_peer_allreduce_i64)
against torch.distributed.all_reduce on the same int64 values -- equality over many back-to-back calls with
rank-dependent delays in between (a fast rank runs ahead of a slow one), then the latency of both on the stream.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/peer_allreduce_check.py
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist

from rubiks_cube_solver_b200 import dist as cdist

rank, local, world = cdist.init_from_env()
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
peer = cdist.PeerCounters(capacity=256)
gen = torch.Generator(device=dev).manual_seed(100 + rank)
bad = 0
for it in range(200):
    n = 1 + (it * 37) % 256
    v = torch.randint(-2 ** 40, 2 ** 40, (n,), dtype=torch.int64, device=dev, generator=gen)
    want = v.clone()
    dist.all_reduce(want, op=dist.ReduceOp.SUM)
    if (it + rank) % 3 == 0:
        torch.cuda._sleep(200000 * (1 + rank))             # this rank reaches the call late
    got = peer.allreduce_(v.clone())
    if (it + rank) % 5 == 0:
        torch.cuda._sleep(100000)
    bad += int((got != want).any())
# three calls in a row without any synchronisation in between (slot double-buffering)
vs = [torch.full((80,), (rank + 1) * (k + 1), dtype=torch.int64, device=dev) for k in range(3)]
for v in vs:
    peer.allreduce_(v)
for k, v in enumerate(vs):
    bad += int((v != (k + 1) * world * (world + 1) // 2).any())
t = torch.tensor([bad], dtype=torch.int64, device=dev)
dist.all_reduce(t)


def timed(fn, reps=200):
    buf = torch.ones(80, dtype=torch.int64, device=dev)
    for _ in range(20):
        fn(buf)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn(buf)
    e1.record()
    torch.cuda.synchronize()
    return cdist.max_over_ranks(e0.elapsed_time(e1) / reps * 1e3, dev)


us_peer = timed(peer.allreduce_)
us_nccl = timed(lambda b: dist.all_reduce(b, op=dist.ReduceOp.SUM))
if rank == 0:
    print(json.dumps({"world": world, "mismatching_calls": int(t[0]), "peer_us_per_call": us_peer, "nccl_us_per_call": us_nccl,
                      "note": "640-byte int64 SUM all-reduce, back-to-back calls on one stream, max over ranks"}))
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if int(t[0]) else 0)
