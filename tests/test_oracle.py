"""The oracle (NumPy, scalar and plain-C restatements) against the golden vectors
that oracle/gen_golden.py produced from the reference itself."""
import hashlib
import os

import numpy as np
import pytest

from oracle import cube_c as C
from oracle import cube_np as O
from oracle import gen_c_tables
from oracle import tables as T
from oracle.gen_golden import ExactValueNet
from oracle.scalar_env import ScalarCubeEnv

from conftest import GOLDEN, golden

SIZES = (2, 3)

# SURVEY.md Appendix A rows (py222 moveDefs for U U' F F' R R')
APPENDIX_A_2 = [
    "2 0 3 1 20 21 6 7 4 5 10 11 12 13 14 15 8 9 18 19 16 17 22 23",
    "1 3 0 2 8 9 6 7 16 17 10 11 12 13 14 15 20 21 18 19 4 5 22 23",
    "0 1 19 17 2 5 3 7 10 8 11 9 6 4 14 15 16 12 18 13 20 21 22 23",
    "0 1 4 6 13 5 12 7 9 11 8 10 17 19 14 15 16 3 18 2 20 21 22 23",
    "0 9 2 11 6 4 7 5 8 13 10 15 12 22 14 20 16 17 18 19 3 21 1 23",
    "0 22 2 20 5 7 4 6 8 1 10 3 12 9 14 11 16 17 18 19 15 21 13 23",
]
APPENDIX_A_PIECEINDS_2 = {50: (0, 0), 54: (0, 1), 13: (0, 2), 28: (1, 0), 42: (1, 1), 8: (1, 2),
                          14: (2, 0), 21: (2, 1), 4: (2, 2), 52: (3, 0), 15: (3, 1), 11: (3, 2),
                          47: (4, 0), 30: (4, 1), 40: (4, 2), 25: (5, 0), 18: (5, 1), 35: (5, 2),
                          23: (6, 0), 57: (6, 1), 37: (6, 2)}


def test_tables_match_reference_dump():
    g = golden("reference_tables.npz")
    assert (T.MOVE_DEFS_3 == g["moveDefs"]).all()
    assert (T.CORNER_DEFS_3 == g["corner_pieceDefs"]).all()
    assert (T.EDGE_DEFS_3 == g["edge_pieceDefs"]).all()
    assert (T.CORNER_INDS_3 == g["corner_pieceInds"]).all()
    assert (T.EDGE_INDS_3 == g["edge_pieceInds"]).all()
    assert (T.SOLVED[3] == g["initState_3"]).all()
    assert list(g["actions_2"]) == T.ACTIONS[2] and list(g["actions_3"]) == T.ACTIONS[3]


def test_tables_2x2_match_appendix_a():
    rows = np.array([[int(x) for x in r.split()] for r in APPENDIX_A_2])
    assert (T.MOVE_DEFS_2 == rows).all()
    want = np.zeros((58, 2), dtype=np.int64)
    for h, po in APPENDIX_A_PIECEINDS_2.items():
        want[h] = po
    assert (T.PIECE_INDS_2 == want).all()


def test_move_table_invariants():
    # permutations, inverse pairs, order 4, fixed centres / fixed DBL cubie
    for size in SIZES:
        M = T.MOVE_DEFS[size]
        S = T.N_STICKERS[size]
        ident = np.arange(S)
        for a in range(T.N_ACTIONS[size]):
            assert sorted(M[a]) == list(ident)
            assert (M[a][M[a ^ 1]] == ident).all()
            r = ident
            for _ in range(4):
                r = r[M[a]]
            assert (r == ident).all()
            assert (M[a] != ident).sum() == (20 if size == 3 else 12)
        fixed = (4, 13, 22, 31, 40, 49) if size == 3 else (14, 18, 23)
        assert (M[:, fixed] == np.array(fixed)).all()


def test_generated_c_header_is_current():
    with open(gen_c_tables.OUT) as f:
        assert f.read() == gen_c_tables.render()


@pytest.mark.parametrize("size", SIZES)
def test_config1_numpy_oracle(size):
    g = golden("config1_%d.npz" % size)
    for seed in (0, 1, 17, 1023):
        assert (O.reference_moves(size, seed, 10) == g["moves"][seed]).all()
    final = O.scramble(size, g["moves"])
    assert (final == g["stickers"]).all()
    assert (O.encode(size, final) == g["onehot"]).all()
    sol = O.is_solved(size, final)
    assert (sol == g["done"]).all()
    assert (O.rewards(sol) == g["reward"]).all()


@pytest.mark.parametrize("size", SIZES)
def test_config1_digests(size):
    want = dict(line.split() for line in open(os.path.join(GOLDEN, "digests.txt")))
    moves = np.stack([O.reference_moves(size, s, 10) for s in range(1024)])
    final = O.scramble(size, moves)
    assert hashlib.sha256(final.tobytes()).hexdigest() == want["config1_%d_stickers_sha256" % size]
    assert hashlib.sha256(O.encode(size, final).tobytes()).hexdigest() == want["config1_%d_onehot_sha256" % size]
    sol = O.is_solved(size, final)
    assert int(sol.sum()) == int(want["config1_%d_solved" % size])
    assert float(O.rewards(sol).sum()) == float(want["config1_%d_reward_sum" % size])


@pytest.mark.parametrize("size", SIZES)
def test_config1_c_oracle(size):
    g = golden("config1_%d.npz" % size)
    final, sol, rew, cnt = C.scramble(size, g["moves"])
    assert (final == g["stickers"]).all()
    assert (sol == g["done"]).all() and (rew == g["reward"]).all() and cnt == int(g["done"].sum())
    assert (C.encode(size, final) == g["onehot"]).all()


@pytest.mark.parametrize("size", SIZES)
def test_walks(size):
    g = golden("walks_%d.npz" % size)
    moves = g["moves"]
    _, trail, flags = O.scramble(size, moves, per_step=True)
    assert (trail == g["stickers"]).all()
    assert (flags == g["done"]).all()
    assert g["done"].any(), "fixture must contain solved steps"
    assert (O.rewards(flags) == g["reward"]).all()
    n, d = moves.shape
    assert (O.encode(size, trail.reshape(n * d, -1)).reshape(g["onehot"].shape) == g["onehot"]).all()
    # plain-C restatement: per-step flags, and single steps on resident states
    _, _, _, _, per = C.scramble(size, moves, per_step=True)
    assert (per == g["done"]).all()
    s = O.solved_states(size, n)
    for k in range(d):
        s, sol, rew = C.step(size, s, moves[:, k])
        assert (s == g["stickers"][:, k]).all() and (sol == g["done"][:, k]).all()
        assert (rew == g["reward"][:, k]).all()


@pytest.mark.parametrize("size", SIZES)
def test_expand(size):
    g = golden("expand_%d.npz" % size)
    children, sol = O.expand(size, g["parents"])
    assert (children == g["children"]).all() and (sol == g["child_solved"]).all()
    n, a = sol.shape
    enc = O.encode(size, children.reshape(n * a, -1)).reshape(g["child_onehot"].shape)
    assert (enc == g["child_onehot"]).all()
    c2, cols, s2 = C.expand(size, g["parents"])
    assert (c2 == g["children"]).all() and (s2 == g["child_solved"]).all()
    assert (cols == g["child_onehot"].argmax(axis=-1)).all()
    assert g["child_solved"].any()


def test_decode_2():
    g = golden("decode_2.npz")
    assert (O.decode_2(g["onehot"]) == g["stickers"]).all()
    assert (O.encode(2, g["stickers"]) == g["onehot"]).all()


@pytest.mark.parametrize("size", SIZES)
def test_scalar_env_matches_golden(size):
    g = golden("config1_%d.npz" % size)
    env = ScalarCubeEnv(size)
    before = np.random.get_state()[1].copy()
    for seed in range(0, 1024, 37):
        obs = env.reset(seed=seed, scramble_count=10)
        assert str(obs.dtype) == str(g["obs_dtype"])
        assert (env.sim_cube == g["stickers"][seed]).all()
        assert (obs == g["onehot"][seed]).all()
    assert (np.random.get_state()[1] == before).all()      # cube_env.py:62,68: RNG state restored
    w = golden("walks_%d.npz" % size)
    env.init_state()
    for k, a in enumerate(w["moves"][0]):
        obs, r, d, info = env.step(int(a))
        assert (obs == w["onehot"][0, k]).all() and r == w["reward"][0, k] and d == w["done"][0, k]
        assert info == {}
    with pytest.raises(IndexError):
        env.step(T.N_ACTIONS[size])
    with pytest.raises(NotImplementedError):
        ScalarCubeEnv(4)


@pytest.mark.parametrize("size", SIZES)
def test_adi_golden(size):
    g = golden("adi_%d.npz" % size)
    moves = g["moves"]
    n, d = moves.shape
    # scalar env reproduces get_random_samples end to end
    env = ScalarCubeEnv(size)
    net = ExactValueNet(T.STATE_DIM[size], T.N_ACTIONS[size])
    assert (net.w.numpy() == g["net_w"]).all()
    buf = []
    saved = np.random.get_state()
    np.random.seed(11)
    env.get_random_samples(buf, net, d, n, float(g["temperature"]))
    np.random.set_state(saved)
    assert (np.array([b["state"] for b in buf]) == g["state"]).all()
    assert [b["target_policy"] for b in buf] == list(g["target_policy"])
    assert [b["target_value"] for b in buf] == list(g["target_value"])
    assert [b["scramble_count"] for b in buf] == list(g["scramble_count"])
    assert np.array_equal(np.array([b["error"] for b in buf]), g["error"])
    # batched oracle: prefixes -> children -> targets
    _, trail, _ = O.scramble(size, moves, per_step=True)
    parents = trail.reshape(n * d, -1)
    assert (O.encode(size, parents) == g["state"]).all()
    children, sol = O.expand(size, parents)
    a = sol.shape[1]
    enc = O.encode(size, children.reshape(n * d * a, -1)).reshape(n * d, a, -1).astype(np.float32)
    cv = (enc * g["net_w"]).sum(axis=2)
    pv = (O.encode(size, parents).reshape(n * d, -1).astype(np.float32) * g["net_w"]).sum(axis=1)
    tv, tp, err = O.adi_targets(cv, sol, pv, np.tile(np.arange(1, d + 1), n), float(g["temperature"]))
    assert (tp == g["target_policy"]).all()
    assert np.array_equal(tv.astype(np.float64), g["target_value"])
    assert np.array_equal(err, g["error"])


def test_undo_is_solved_and_hypothesis_properties():
    rng = np.random.RandomState(0)
    for size in SIZES:
        A = T.N_ACTIONS[size]
        fwd = rng.randint(A, size=(256, 17))
        both = np.concatenate((fwd, fwd[:, ::-1] ^ 1), axis=1)
        final = O.scramble(size, both)
        assert O.is_solved(size, final).all()
        assert (final == O.solved_states(size, 256)).all()
        # expand[a] == step(a)
        s = O.scramble(size, fwd)
        children, _ = O.expand(size, s)
        for a in range(A):
            assert (children[:, a] == O.apply_moves(size, s, np.full(256, a))).all()


def test_reset_sequence_equals_one_big_draw():
    # legacy RandomState: count successive randint(A, size=d) calls == one randint(A, size=(count,d))
    for A in (6, 12):
        r1, r2 = np.random.RandomState(5), np.random.RandomState(5)
        a = np.stack([r1.randint(A, size=30) for _ in range(20)])
        b = r2.randint(A, size=(20, 30))
        assert (a == b).all()


@pytest.mark.parametrize("size", (2, 3))
def test_mcts_oracle_matches_reference_golden(size):
    """oracle/mcts_ref.py against the reference's own mcts.py (golden vectors generated by running it,
    oracle/gen_golden.py::_mcts): action lists, simulations used, node counts, root statistics."""
    import random
    import torch
    from oracle import mcts_ref
    from oracle.gen_golden import ExactSearchNet, MCTS_CFG
    from oracle.scalar_env import ScalarCubeEnv
    g = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mcts_%d.npz" % size))
    env = ScalarCubeEnv(size)
    net = ExactSearchNet(env.state_dim, env.action_dim)
    cfg = MCTS_CFG["mcts"]
    for i, (seed, depth) in enumerate(g["cases"]):
        obs = env.reset(seed=int(seed), scramble_count=int(depth))
        with torch.no_grad():
            acts, used, tree = mcts_ref.solve(net.predict, env, obs, cfg["numMCTSSim"], random.Random(1000 + int(seed)),
                                              loss_constant=cfg["virtual_loss_const"], cpuct=cfg["cpuct"],
                                              value_min=cfg["value_min"])
        root = tree.nodes[tree.key(obs)]
        assert (acts or []) == g["actions"][i][:g["n_actions"][i]].tolist()
        assert used == g["n_sims"][i] and len(tree.nodes) == g["n_nodes"][i]
        assert [int(v) for v in root[3]] == g["root_N"][i].tolist()
        assert [float(v) for v in root[2]] == g["root_W"][i].tolist()
        assert [int(v) for v in root[4]] == g["root_L"][i].tolist()


def test_exact_3x3_encoding_spec_is_a_bijection():
    """oracle.cube_np.encode_exact / decode_exact (the spec of the opt-in CUBE_ENCODING_EXACT): every corner triple
    of a reachable state is assigned, pieces form a permutation, twists / flips obey the cube's parities,
    decode inverts encode, and the reference encoding provably is NOT injective on the same states."""
    from oracle import cube_np as O
    rng = np.random.RandomState(3)
    s = O.scramble(3, rng.randint(12, size=(30000, 45)))
    cols = O.onehot_columns_exact(s)
    assert (cols >= 0).all()
    assert (np.sort(cols[:, :8] // 3, axis=1) == np.arange(8)).all() and (np.sort(cols[:, 8:] // 2, axis=1) == np.arange(12)).all()
    assert ((cols[:, :8] % 3).sum(1) % 3 == 0).all() and ((cols[:, 8:] % 2).sum(1) % 2 == 0).all()
    enc = O.encode_exact(s)
    assert (enc.sum(axis=2) == 1).all() and (O.decode_exact(enc) == s).all()
    assert (O.onehot_columns_exact(O.solved_states(3, 1))[0] == np.concatenate((3 * np.arange(8), 2 * np.arange(12)))).all()
    ref = O.encode(3, s)
    assert (ref[:, 8:] == enc[:, 8:]).all()
    # the shipped table maps different corner configurations to the same observation
    corner_cfg = np.unique(s[:, O.CORNER_DEFS_EXACT.ravel()], axis=0).shape[0]
    assert np.unique(ref[:, :8].argmax(2), axis=0).shape[0] < corner_cfg == np.unique(cols[:, :8], axis=0).shape[0]
