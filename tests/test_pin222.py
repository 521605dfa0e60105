"""The reference's own trained 2x2x2 checkpoint as a pin of the restated py222 semantics.

`assets/py222.py` is absent from the reference tree; its moves and its getOP tables are restated
(SURVEY.md Appendix A).  `tests/golden/pin222.npz` (oracle/gen_pin222.py) holds the weights of the
reference's `pretrained/222model.pt` -- trained on the REAL py222 -- and the episodes of the reference's own
greedy loop (cube_env.py + model.py, train.py:167-198 / test.py:103-158) run on the restatement: 100 % solved
at scramble depths 1-5, 96.5 % at 8; with a wrong encoding (orientation labels swapped) 25 % / 8 % / 1 %.
A net trained on py222 can only solve cubes through tables that ARE py222's, so these episodes pin the
restatement statistically; the tests below hold the oracle and the CUDA path to the same episodes."""
import os

import numpy as np
import pytest
import torch

from oracle import cube_np as O
from oracle import tables as T

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "pin222.npz")


class DeepCube222(torch.nn.Module):
    """Layer shapes of model.py:7-29 for the shipped checkpoint (147-512-128-{64-6, 64-1}, ELU)."""

    def __init__(self, g):
        super().__init__()
        nn = torch.nn
        self.encoder_net = nn.Sequential(nn.Flatten(), nn.Linear(147, 512), nn.ELU(), nn.Linear(512, 128), nn.ELU())
        self.policy_net = nn.Sequential(nn.Linear(128, 64), nn.ELU(), nn.Linear(64, 6))
        self.value_net = nn.Sequential(nn.Linear(128, 64), nn.ELU(), nn.Linear(64, 1))
        self.load_state_dict({k[2:]: torch.from_numpy(g[k]) for k in g.files if k.startswith("w:")})
        self.eval()

    def forward(self, x):
        h = self.encoder_net(x)
        return self.value_net(h), self.policy_net(h)


def _scrambles(g):
    """reset(seed, depth) of every (depth, seed) episode (cube_env.py:56-69), row-major over (depth, seed)."""
    rows = []
    for d in g["depths"]:
        for s in g["seeds"]:
            rows.append(O.scramble(2, np.random.RandomState(int(s)).randint(6, size=(1, int(d))))[0])
    return np.stack(rows)


def _greedy_oracle(model, states, max_t, mask, action_map=None, encode=None):
    """The reference loop for a batch, on the CPU oracle: returns (solved, steps, actions)."""
    n = states.shape[0]
    states = states.copy()
    done = np.zeros(n, dtype=bool)
    steps = np.zeros(n, dtype=np.int64)
    acts = np.full((n, max_t), -1, dtype=np.int8)
    pre = np.full(n, -1, dtype=np.int64)
    for t in range(1, max_t + 1):
        live = np.flatnonzero(~done)
        if live.size == 0:
            break
        obs = (encode or O.encode)(2, states[live], dtype=np.float32)
        with torch.no_grad():
            logits = model(torch.from_numpy(obs))[1].numpy()
        order = np.argsort(-logits, axis=1, kind="stable")
        a = order[:, 0]
        if mask:                                             # model.get_action with pre_action (model.py:47-76)
            invalid = np.where(pre[live] >= 0, pre[live] ^ 1, -1)
            a = np.where(a == invalid, order[:, 1], a)
            pre[live] = a
        acts[live, t - 1] = a
        states[live] = O.apply_moves(2, states[live], a if action_map is None else action_map[a])
        ok = O.is_solved(2, states[live])
        steps[live[ok]] = t
        done[live[ok]] = True
    return done, steps, acts


@pytest.fixture(scope="module")
def pin():
    g = np.load(GOLDEN)
    return g, DeepCube222(g), _scrambles(g)


@pytest.mark.parametrize("mask", (False, True))
def test_oracle_reproduces_the_reference_checkpoints_episodes(pin, mask):
    g, model, start = pin
    tag = "mask" if mask else "plain"
    solved, steps, acts = _greedy_oracle(model, start, int(g["max_timesteps"]), mask)
    shape = g["solved_" + tag].shape
    assert (solved.reshape(shape) == g["solved_" + tag]).all()
    assert (steps.reshape(shape) == g["steps_" + tag]).all()
    assert (acts.reshape(shape + (-1,)) == g["actions_" + tag]).all()
    rate = g["solved_" + tag].mean(axis=1)                    # per depth 1, 2, 3, 5, 8, 12
    assert (rate[:4] == 1.0).all() and rate[4] >= 0.95 and rate[5] >= 0.8


def test_pin_is_sensitive_to_wrong_semantics(pin):
    """The checkpoint stops solving as soon as the restated tables are perturbed: a wrong encoding (the
    control recorded with the reference's own loop), or moves whose direction is exchanged."""
    g, model, start = pin
    assert (g["control_solve_rate"] <= 0.30).all() and g["control_solve_rate"][2:].max() <= 0.15
    n = len(g["seeds"])
    deep = start[2 * n:5 * n]                                 # depths 3, 5, 8
    flipped, _, _ = _greedy_oracle(model, deep, int(g["max_timesteps"]), False, action_map=np.arange(6) ^ 1)
    assert flipped.mean() <= 0.10                              # every turn goes the other way
    swapped, _, _ = _greedy_oracle(model, deep, int(g["max_timesteps"]), False, action_map=np.array([2, 3, 0, 1, 4, 5]))
    assert swapped.mean() <= 0.30                              # two faces exchanged

    def rolled(size, s, dtype):                               # cubelet rows rotated by one: a wrong getOP table
        return np.roll(O.encode(size, s, dtype=dtype), 1, axis=1)
    wrong, _, _ = _greedy_oracle(model, deep, int(g["max_timesteps"]), False, encode=rolled)
    assert wrong.mean() <= 0.30


@pytest.mark.gpu
@pytest.mark.parametrize("mask", (False, True))
def test_cuda_path_reproduces_the_reference_checkpoints_episodes(pin, mask):
    """The same episodes through the CUDA path (fused scramble, encode, batched step: rollout.greedy_solve)."""
    from rubiks_cube_solver_b200 import rollout
    g, model, _ = pin
    tag = "mask" if mask else "plain"
    moves = rollout.reference_scrambles(2, [int(s) for s in g["seeds"]], [int(d) for d in g["depths"]])
    res = rollout.greedy_solve(model.cuda(), 2, moves, max_timesteps=int(g["max_timesteps"]), mask_inverse=mask)
    solved = res["solved"].cpu().numpy().reshape(g["solved_" + tag].shape)
    steps = res["steps"].cpu().numpy().reshape(solved.shape)
    # fp32 GEMMs on the GPU may order two nearly equal logits differently: allow a handful of episodes
    assert (solved != g["solved_" + tag]).mean() <= 0.005 and (steps != g["steps_" + tag]).mean() <= 0.01
    rate = solved.mean(axis=1)
    assert (rate[:4] == 1.0).all() and rate[4] >= 0.95 and rate[5] >= 0.8
    model.cpu()


@pytest.mark.gpu
def test_batched_mcts_with_the_reference_checkpoint(pin):
    """BASELINE config 5's consumer end to end: the reference's own mcts.py with the checkpoint as its net
    (oracle/gen_pin222.py: 50 simulations, 40 seeds at depths 5 / 8 / 12) against BatchedMCTS on the GPU with
    the same weights and the same per-tree random seeds.  The scores are float32 sums of net outputs, and the
    GPU's GEMM may round a logit differently from the CPU's, so a few trees may part ways: every returned
    action list must solve its cube, and the solved flags must agree for at least 90 % of the trees."""
    from rubiks_cube_solver_b200 import mcts_batch
    g, model, _ = pin
    cases = g["mcts_cases"]
    roots = np.stack([O.scramble(2, np.random.RandomState(int(s)).randint(6, size=(1, int(d))))[0] for s, d in cases])
    search = mcts_batch.BatchedMCTS(model.cuda(), 2, num_sim=50, cpuct=1.0, virtual_loss_const=150, value_min=-10.0)
    out = search.run(torch.from_numpy(roots).cuda(), seeds=[1000 + int(s) for s, _ in cases])
    solved = out["solved"].cpu().numpy()
    acts, n_act = out["actions"].cpu().numpy(), out["n_actions"].cpu().numpy()
    for i in np.flatnonzero(solved):
        end = O.scramble(2, acts[i, :n_act[i]][None].astype(np.int64), init=roots[i:i + 1])
        assert O.is_solved(2, end)[0]
    assert (solved == g["mcts_solved"]).mean() >= 0.9
    same = solved & g["mcts_solved"]
    assert (out["n_sims"].cpu().numpy()[same] == g["mcts_n_sims"][same]).mean() >= 0.9
    model.cpu()
