"""Launch each hot kernel a few times at its BASELINE size (for ncu captures and quick timing).

    python tools/run_kernels.py [scramble3|scramble3_d32|scramble2|step3|step2|expand3|leaf2|all] [--iters N]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch

from rubiks_cube_solver_b200 import ops


def main():
    which = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "all"
    iters = int(sys.argv[sys.argv.index("--iters") + 1]) if "--iters" in sys.argv else 5
    dev = torch.device("cuda", 0)
    gen = torch.Generator(device=dev).manual_seed(1234)

    def timed(name, fn, units):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        print("%-10s %8.4f ms  %.4g units/s" % (name, ms, units / ms * 1e3), flush=True)

    if which in ("scramble3", "all"):
        n, d = 8 * 2 ** 20, 30
        moves = torch.randint(0, 12, (n, d), dtype=torch.uint8, device=dev, generator=gen)
        st = torch.empty((n, 54), dtype=torch.uint8, device=dev)
        so = torch.empty(n, dtype=torch.uint8, device=dev)
        rw = torch.empty(n, dtype=torch.float32, device=dev)
        timed("scramble3", lambda: ops.scramble(3, moves, out=st, solved=so, reward=rw), n * d)
    if which == "scramble3_d32":                   # a depth whose move tile is staged swizzled
        n, d = 8 * 2 ** 20, 32
        moves = torch.randint(0, 12, (n, d), dtype=torch.uint8, device=dev, generator=gen)
        st = torch.empty((n, 54), dtype=torch.uint8, device=dev)
        so = torch.empty(n, dtype=torch.uint8, device=dev)
        rw = torch.empty(n, dtype=torch.float32, device=dev)
        timed("scramble3_d32", lambda: ops.scramble(3, moves, out=st, solved=so, reward=rw), n * d)
    if which == "seeds3":                          # K0: identically seeded move draws, 8 Mi seeds x depth 30
        import ctypes
        from rubiks_cube_solver_b200 import _lib
        n, d = 8 * 2 ** 20, 30
        seeds = torch.arange(n, dtype=torch.int32, device=dev)
        mv = torch.empty((n, d), dtype=torch.uint8, device=dev)
        lib = _lib.load()
        stream = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        timed("seeds3", lambda: lib.cube_moves_from_seeds(3, ctypes.c_void_p(seeds.data_ptr()), n, d,
                                                          ctypes.c_void_p(mv.data_ptr()), None, stream), n * d)
        want = ops.moves_from_seeds(3, list(range(1000)), d)
        assert bool((mv[:1000] == want).all())
    if which in ("scramble2", "all"):
        n, d = 16 * 2 ** 20, 20
        moves = torch.randint(0, 6, (n, d), dtype=torch.uint8, device=dev, generator=gen)
        st = torch.empty((n, 24), dtype=torch.uint8, device=dev)
        so = torch.empty(n, dtype=torch.uint8, device=dev)
        rw = torch.empty(n, dtype=torch.float32, device=dev)
        timed("scramble2", lambda: ops.scramble(2, moves, out=st, solved=so, reward=rw), n * d)
    for size, n, key in ((3, 8 * 2 ** 20, "step3"), (2, 16 * 2 ** 20, "step2")):
        if which in (key, "all"):
            states = ops.solved_states(size, n, dev)
            act = torch.randint(0, ops.N_ACTIONS[size], (n,), dtype=torch.uint8, device=dev, generator=gen)
            so = torch.empty(n, dtype=torch.uint8, device=dev)
            rw = torch.empty(n, dtype=torch.float32, device=dev)
            timed(key, lambda: ops.step(size, states, act, solved=so, reward=rw), n)
    if which in ("expand3", "all"):
        n = 2 ** 20
        parents, _, _ = ops.scramble(3, torch.randint(0, 12, (n, 30), dtype=torch.uint8, device=dev, generator=gen),
                                     want_flags=False)
        child = torch.empty((n, 12, 20, 24), dtype=torch.bfloat16, device=dev)
        timed("expand3", lambda: ops.expand(3, parents, dtype=torch.bfloat16, child_onehot=child), n)
    if which in ("leaf2", "all"):
        n = 2 ** 20
        leaves, _, _ = ops.scramble(2, torch.randint(0, 6, (n, 11), dtype=torch.uint8, device=dev, generator=gen),
                                    want_flags=False)
        timed("leaf2", lambda: ops.expand(2, leaves, dtype=torch.bfloat16, want_children=True,
                                          want_child_onehot=False, want_parent_onehot=True), n)
    for size, n, key in ((3, 4 * 2 ** 20, "encode3"), (2, 8 * 2 ** 20, "encode2")):
        if which in (key, "all"):
            a = ops.N_ACTIONS[size]
            st, _, _ = ops.scramble(size, torch.randint(0, a, (n, 15), dtype=torch.uint8, device=dev, generator=gen),
                                    want_flags=False)
            obs = torch.empty((n,) + ops.STATE_DIM[size], dtype=torch.bfloat16, device=dev)
            timed(key, lambda: ops.encode(size, st, dtype=torch.bfloat16, out=obs), n)
            d = ops.STATE_DIM[size][0] * ops.STATE_DIM[size][1]
            print("   %s algorithmic bytes/row %d" % (key, ops.N_STICKERS[size] + 2 * d))
    if which in ("expand2", "all"):
        n = 4 * 2 ** 20
        parents, _, _ = ops.scramble(2, torch.randint(0, 6, (n, 14), dtype=torch.uint8, device=dev, generator=gen),
                                     want_flags=False)
        child = torch.empty((n, 6, 7, 21), dtype=torch.bfloat16, device=dev)
        timed("expand2", lambda: ops.expand(2, parents, dtype=torch.bfloat16, child_onehot=child), n)
        print("   expand2 algorithmic bytes/parent %d" % (24 + 6 * (294 + 5)))
    if which in ("prefixes3", "small"):            # K1x: every prefix of 139 810 scrambles x depth 30 (config 4's parents)
        n, d = 139810, 30
        moves = torch.randint(0, 12, (n, d), dtype=torch.uint8, device=dev, generator=gen)
        out = torch.empty((n, d, 54), dtype=torch.uint8, device=dev)
        timed("prefixes3", lambda: ops.scramble_prefixes(3, moves, out=out), n * d)
        print("   prefixes3 algorithmic bytes/cube %d" % (d + d * 54))
    if which in ("adi_targets", "small"):          # K4: target assembly for 4 Mi parents
        p = 4 * 2 ** 20
        cv = torch.randn((p, 12), device=dev, generator=gen)
        so = (torch.rand((p, 12), device=dev, generator=gen) < 0.01).to(torch.uint8)
        pv = torch.randn(p, device=dev, generator=gen)
        k = torch.randint(1, 31, (p,), device=dev, generator=gen).to(torch.int32)
        timed("adi_targets", lambda: ops.adi_targets(3, cv, so, pv, k, 1.0), p)
        print("   adi_targets algorithmic bytes/parent %d (incl. the wrapper's int64 policy copy)" % (48 + 12 + 4 + 4 + 4 + 4 + 8))
    if which in ("decode2", "small"):              # 2x2x2 one-hot -> stickers, 4 Mi rows of bf16
        n = 4 * 2 ** 20
        st, _, _ = ops.scramble(2, torch.randint(0, 6, (n, 14), dtype=torch.uint8, device=dev, generator=gen), want_flags=False)
        oh = ops.encode(2, st, dtype=torch.bfloat16)
        timed("decode2", lambda: ops.decode(2, oh), n)
        assert bool((ops.decode(2, oh) == st).all())
        print("   decode2 algorithmic bytes/row %d" % (294 + 24))
    if which in ("decode3x", "small"):             # 3x3x3 EXACT one-hot -> stickers, 4 Mi rows of bf16
        n = 4 * 2 ** 20
        st, _, _ = ops.scramble(3, torch.randint(0, 12, (n, 25), dtype=torch.uint8, device=dev, generator=gen), want_flags=False)
        oh = ops.encode(3, st, dtype=torch.bfloat16, encoding="exact")
        timed("decode3x", lambda: ops.decode(3, oh, encoding="exact"), n)
        assert bool((ops.decode(3, oh, encoding="exact") == st).all())
        print("   decode3x algorithmic bytes/row %d" % (960 + 54))
    if which in ("scramble3_d1000", "small"):      # K1p sliced: 1 Mi instances x depth 1000 (test.py:44 scrambles that deep)
        n, d = 2 ** 20, 1000
        moves = torch.randint(0, 12, (n, d), dtype=torch.uint8, device=dev, generator=gen)
        st = torch.empty((n, 54), dtype=torch.uint8, device=dev)
        so = torch.empty(n, dtype=torch.uint8, device=dev)
        rw = torch.empty(n, dtype=torch.float32, device=dev)
        timed("scramble3_d1000", lambda: ops.scramble(3, moves, out=st, solved=so, reward=rw), n * d)
        print("   scramble3_d1000 algorithmic bytes/instance %d" % (d + 54 + 5))
    if which in ("scramble_step3", "small"):       # fused scramble + step (configs[2] read literally), 8 Mi x depth 30 + 1
        n, d = 8 * 2 ** 20, 30
        moves = torch.randint(0, 12, (n, d), dtype=torch.uint8, device=dev, generator=gen)
        act = torch.randint(0, 12, (n,), dtype=torch.uint8, device=dev, generator=gen)
        st = torch.empty((n, 54), dtype=torch.uint8, device=dev)
        so = torch.empty(n, dtype=torch.uint8, device=dev)
        rw = torch.empty(n, dtype=torch.float32, device=dev)
        timed("scramble_step3", lambda: ops.scramble_step(3, moves, act, out=st, solved=so, reward=rw), n * (d + 1))
        print("   scramble_step3 algorithmic bytes/instance %d" % (d + 1 + 54 + 5))
    if which == "expand2_f32":                     # what the drop-in env / ADI feed the reference's float32 net
        n = 2 * 2 ** 20
        parents, _, _ = ops.scramble(2, torch.randint(0, 6, (n, 14), dtype=torch.uint8, device=dev, generator=gen),
                                     want_flags=False)
        child = torch.empty((n, 6, 7, 21), dtype=torch.float32, device=dev)
        timed("expand2_f32", lambda: ops.expand(2, parents, dtype=torch.float32, child_onehot=child), n)
        print("   expand2_f32 algorithmic bytes/parent %d" % (24 + 6 * (588 + 5)))


if __name__ == "__main__":
    main()
