"""Generate tests/golden/*.npz from the reference itself (TEST INFRASTRUCTURE ONLY).

Run in the build container (the only place /root/reference exists):

    python -m oracle.gen_golden

Every array written here is produced by the reference's own code
(`gym-cube/gym_cube/envs/cube_env.py`, `assets/py333.py`, `env.py`) imported
from /root/reference under the test-only shims of oracle/ref_harness.py.  For
cube_size=2 the reference's cube_env.py runs on top of the stand-in
`assets/py222.py` (oracle/ref_shims), because the real py222 is not in the
reference tree -- see oracle/__init__.py ("parity unpinned" at that boundary).
The fixtures travel to the GPU box; /root/reference does not.
"""
import hashlib
import os

import numpy as np
import torch

from . import ref_harness

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


class ExactValueNet(torch.nn.Module):
    """Deterministic stand-in for DeepCube with the same call surface
    (model.py:31-45: returns (value[B,1], policy[B,A])).  Weights are small
    multiples of 1/8 so every value is exact in fp32 on any machine."""

    def __init__(self, state_dim, action_dim, seed=7):
        super().__init__()
        r = np.random.RandomState(seed)
        d = state_dim[0] * state_dim[1]
        self.w = torch.tensor(r.randint(-8, 9, size=(d,)).astype(np.float32) / 8.0)
        self.action_dim = action_dim

    def forward(self, x):
        if x.dim() == 2:
            x = x.unsqueeze(0)
        v = (x.reshape(x.shape[0], -1).float() * self.w).sum(dim=1, keepdim=True)
        return v, torch.zeros(x.shape[0], self.action_dim)


class ExactSearchNet(torch.nn.Module):
    """DeepCube-shaped net for the MCTS vectors whose value AND softmax policy are bit-identical on
    any device: the value is a sum of multiples of 1/8; the logits are 0 or -inf (chosen by small
    integer hashes of the observation), so softmax is 1/m on m actions and exactly 0 on the rest.
    `predict` restates model.py:78-91."""

    def __init__(self, state_dim, action_dim, seed=5):
        super().__init__()
        r = np.random.RandomState(seed)
        d = state_dim[0] * state_dim[1]
        self.register_buffer("wv", torch.tensor(r.randint(-8, 9, size=(d,)).astype(np.float32) / 8.0))
        self.register_buffer("wm", torch.tensor(r.randint(0, 4, size=(d, action_dim)).astype(np.float32)))
        self.action_dim = action_dim

    def forward(self, x):
        if x.dim() == 2:
            x = x.unsqueeze(0)
        flat = x.reshape(x.shape[0], -1).float()
        v = (flat * self.wv).sum(dim=1, keepdim=True)
        h = flat @ self.wm                                          # small integers: exact
        banned = torch.remainder(h, 3.0) == 0
        banned[:, 0] = False
        logits = torch.where(banned, torch.full_like(h, float("-inf")), torch.zeros_like(h))
        return v, logits

    def predict(self, x):
        x = torch.tensor(x).float().detach()
        value, policy = self.forward(x)
        policy = torch.nn.functional.softmax(policy, dim=-1)
        return value.numpy()[0], policy.numpy()[0]


MCTS_CFG = {"mcts": {"numMCTSSim": 50, "cpuct": 1.0, "virtual_loss_const": 150, "value_min": -10.0}}


def mcts_cases(size):
    """(seed, scramble depth) of the golden MCTS searches."""
    return [(s, 1 + s % (5 if size == 3 else 7)) for s in range(16)]


def _mcts(R, size):
    """The reference's own MCTS (mcts.py) on its own env: test.py's MCTS branch for one time step
    (test.py:139-147), `random` seeded per cube with 1000 + seed."""
    import importlib
    import random
    mcts_mod = importlib.import_module("mcts")
    env = R.make_env(size)
    net = ExactSearchNet(env.state_dim, env.action_dim)
    cfg = dict(MCTS_CFG, test={"cube_size": size})
    n_sim = cfg["mcts"]["numMCTSSim"]
    A = env.action_dim
    rows = dict(actions=[], n_actions=[], n_sims=[], n_nodes=[], root_N=[], root_W=[], root_L=[])
    for seed, depth in mcts_cases(size):
        state = env.reset(seed=seed, scramble_count=depth)
        random.seed(1000 + seed)
        tree = mcts_mod.MCTS(net, cfg)
        result, used = None, n_sim
        with torch.no_grad():
            for k in range(n_sim):
                result = tree.train(state, env)
                if result is not None:
                    used = k + 1
                    break
        root = tree.children_and_data[np.array2string(state)]
        acts = list(result) if result is not None else []
        rows["actions"].append(acts + [-1] * (n_sim + 1 - len(acts)))
        rows["n_actions"].append(len(acts))
        rows["n_sims"].append(used)
        rows["n_nodes"].append(len(tree.children_and_data))
        rows["root_N"].append([int(v) for v in root[tree.n_of_v_i]])
        rows["root_W"].append([float(np.asarray(v).reshape(-1)[0]) for v in root[tree.s_i]])
        rows["root_L"].append([int(v) for v in root[tree.v_l_i]])
    out = {k: np.array(v) for k, v in rows.items()}
    out["cases"] = np.array(mcts_cases(size))
    return out


def _config1(R, size, n_seeds=1024, depth=10):
    env = R.make_env(size)
    moves, stickers, onehot, reward, done = [], [], [], [], []
    for seed in range(n_seeds):
        moves.append(np.random.RandomState(seed).randint(env.action_dim, size=depth))
        obs = env.reset(seed=seed, scramble_count=depth)
        # the final step's reward/done are not returned by reset(); recompute them the
        # way step() does (cube_env.py:89-104) from the reference's own solved test
        if size == 3:
            solved = bool(R.py333.isSolved_3(env.sim_cube))
        else:
            solved = bool(R.cube_env.isSolved(env.sim_cube))
        stickers.append(env.sim_cube.copy())
        onehot.append(np.asarray(obs).copy())
        done.append(solved)
        reward.append(1.0 if solved else -1.0)
    return dict(moves=np.array(moves, dtype=np.uint8), stickers=np.array(stickers, dtype=np.uint8),
                onehot=np.array(onehot).astype(np.uint8), reward=np.array(reward, dtype=np.float32),
                done=np.array(done, dtype=bool), obs_dtype=str(np.asarray(obs).dtype))


def _walks(R, size, n_walks=48, depth=30, seed=2024):
    """step() trajectories incl. walks that return to solved (reward +1, done)."""
    env = R.make_env(size)
    rng = np.random.RandomState(seed)
    A = env.action_dim
    seqs = rng.randint(A, size=(n_walks, depth))
    half = depth // 2
    seqs[: n_walks // 3, half:2 * half] = seqs[: n_walks // 3, :half][:, ::-1] ^ 1
    stickers = np.zeros((n_walks, depth, len(env.sim_cube)), dtype=np.uint8)
    onehot = np.zeros((n_walks, depth) + tuple(env.state_dim), dtype=np.uint8)
    reward = np.zeros((n_walks, depth), dtype=np.float32)
    done = np.zeros((n_walks, depth), dtype=bool)
    for w in range(n_walks):
        env.init_state()
        for k in range(depth):
            obs, r, d, info = env.step(int(seqs[w, k]))
            assert info == {}
            stickers[w, k] = env.sim_cube
            onehot[w, k] = obs
            reward[w, k] = r
            done[w, k] = d
    return dict(moves=seqs.astype(np.uint8), stickers=stickers, onehot=onehot, reward=reward, done=done)


def _adi(R, size, n_cubes=6, depth=8, seed=11, temperature=1.0):
    """get_random_samples + get_target_value with an exactly-representable value net."""
    env = R.make_env(size)
    net = ExactValueNet(env.state_dim, env.action_dim)
    buf = []
    saved = np.random.get_state()
    np.random.seed(seed)
    env.get_random_samples(buf, net, depth, n_cubes, temperature)
    np.random.set_state(saved)
    moves = np.random.RandomState(seed).randint(env.action_dim, size=(n_cubes, depth))
    return dict(
        moves=moves.astype(np.uint8), net_w=net.w.numpy(), temperature=np.float64(temperature),
        state=np.array([s['state'] for s in buf]).astype(np.uint8),
        target_value=np.array([s['target_value'] for s in buf], dtype=np.float64),
        target_policy=np.array([s['target_policy'] for s in buf], dtype=np.int64),
        scramble_count=np.array([s['scramble_count'] for s in buf], dtype=np.int64),
        error=np.array([s['error'] for s in buf], dtype=np.float64))


def _expand(R, size, n=96, seed=5):
    """All children of assorted states through the reference's move + encode + solved."""
    env = R.make_env(size)
    rng = np.random.RandomState(seed)
    A = env.action_dim
    parents, children, child_onehot, child_solved = [], [], [], []
    for i in range(n):
        env.init_state()
        for a in rng.randint(A, size=(i % 7) + 1):
            env.step(int(a))
        parent = env.sim_cube.copy()
        parents.append(parent)
        row_c, row_o, row_s = [], [], []
        for a in range(A):
            name = env.action_to_sim_action[size][a]
            if size == 3:
                c = R.py333.doMove_3(parent, name)
                s = bool(R.py333.isSolved_3(c))
            else:
                c = R.cube_env.doMove(parent, name)
                s = bool(R.cube_env.isSolved(c))
            row_c.append(c)
            row_o.append(env.sim_state_to_state(c))
            row_s.append(s)
        children.append(row_c)
        child_onehot.append(row_o)
        child_solved.append(row_s)
    return dict(parents=np.array(parents, dtype=np.uint8), children=np.array(children, dtype=np.uint8),
                child_onehot=np.array(child_onehot).astype(np.uint8),
                child_solved=np.array(child_solved, dtype=bool))


def main():
    R = ref_harness.load_reference()
    os.makedirs(OUT, exist_ok=True)
    p = R.py333
    np.savez_compressed(
        os.path.join(OUT, "reference_tables.npz"),
        moveDefs=p.moveDefs.astype(np.uint8), corner_pieceDefs=p.corner_pieceDefs.astype(np.uint8),
        edge_pieceDefs=p.edge_pieceDefs.astype(np.uint8), corner_pieceInds=p.corner_pieceInds.astype(np.uint8),
        edge_pieceInds=p.edge_pieceInds.astype(np.uint8), initState_3=p.initState_3().astype(np.uint8),
        actions_2=np.array(R.make_env(2).action_to_sim_action[2]),
        actions_3=np.array(R.make_env(3).action_to_sim_action[3]))
    digests = {}
    for size in (2, 3):
        c1 = _config1(R, size)
        np.savez_compressed(os.path.join(OUT, "config1_%d.npz" % size), **c1)
        digests["config1_%d_stickers_sha256" % size] = hashlib.sha256(c1["stickers"].tobytes()).hexdigest()
        digests["config1_%d_onehot_sha256" % size] = hashlib.sha256(c1["onehot"].tobytes()).hexdigest()
        digests["config1_%d_solved" % size] = int(c1["done"].sum())
        digests["config1_%d_reward_sum" % size] = float(c1["reward"].sum())
        np.savez_compressed(os.path.join(OUT, "walks_%d.npz" % size), **_walks(R, size))
        np.savez_compressed(os.path.join(OUT, "adi_%d.npz" % size), **_adi(R, size))
        np.savez_compressed(os.path.join(OUT, "expand_%d.npz" % size), **_expand(R, size))
        np.savez_compressed(os.path.join(OUT, "mcts_%d.npz" % size), **_mcts(R, size))
    # 2x2x2 decode (state_to_sim_state, cube_env.py:154-175)
    env = R.make_env(2)
    rng = np.random.RandomState(3)
    obs, dec = [], []
    for _ in range(64):
        env.init_state()
        for a in rng.randint(6, size=12):
            o, _, _, _ = env.step(int(a))
        obs.append(o.copy())
        dec.append(env.state_to_sim_state(o))
        assert (dec[-1] == env.sim_cube).all()
    np.savez_compressed(os.path.join(OUT, "decode_2.npz"), onehot=np.array(obs).astype(np.uint8),
                        stickers=np.array(dec, dtype=np.uint8))
    with open(os.path.join(OUT, "digests.txt"), "w") as f:
        for k in sorted(digests):
            f.write("%s %s\n" % (k, digests[k]))
    for k in sorted(digests):
        print(k, digests[k])


if __name__ == "__main__":
    main()
