"""Statistical pin of the 2x2x2 semantics by the reference's OWN trained checkpoint (TEST INFRASTRUCTURE ONLY).

`assets/py222.py` is third-party and absent from the reference tree, so the 2x2x2 arithmetic (move rows,
getOP tables) is a restatement that no test of the reference pins.  The reference does
ship `pretrained/222model.pt`, a value / policy net trained ON the real py222: if the restated moves or the
restated encoding were wrong, the net would see garbage and would not solve cubes.  This script runs the
reference's own `cube_env.py` + `model.py` + checkpoint (through the shims of oracle/ref_harness.py: the only
stand-in is the restated py222) over the greedy loop of train.py:167-198 / test.py:103-158 and writes

    tests/golden/pin222.npz
        weights of the checkpoint's `model_state_dict` (fp32, 158 407 parameters), its epoch,
        depths, seeds, max_timesteps, and per (depth, seed) episode of the reference loop:
        solved flag, time step of the first `done`, the action list (-1 padded), with and without
        the `mask` option of test.py (model.get_action(state, pre_action), model.py:47-76);
        the same episodes under a CONTROL encoding (orientation labels 1 <-> 2 swapped in the one-hot):
        solve rates only.

Run in the build container (needs /root/reference):  python oracle/gen_pin222.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import ref_harness  # noqa: E402

DEPTHS = (1, 2, 3, 5, 8, 12)
SEEDS = tuple(range(200))
MAX_T = 50


def swap_orientation_labels(obs):
    """one-hot [7, 21] with column 3 * position + ori: exchange ori 1 and 2 (the control encoding)."""
    out = np.array(obs, copy=True)
    v = out.reshape(7, 7, 3)
    v[:, :, [1, 2]] = v[:, :, [2, 1]]
    return v.reshape(7, 21)


def main():
    import torch
    ref = ref_harness.load_reference()
    path = os.path.join(ref_harness.REFERENCE_ROOT, "pretrained", "222model.pt")
    ck = torch.load(path, map_location="cpu", weights_only=False)           # pickles utils.SharedAdam (test.py:55-60)
    hidden = [ck["optimizer_state_dict"]["state"][x]["exp_avg"].size(0) for x in (1, 3, 5)]     # test.py:57
    model = ref.model.DeepCube([7, 21], 6, hidden)
    model.load_state_dict(ck["model_state_dict"])
    model.eval()
    env = ref.make_env(2)

    def episode(seed, depth, mask, control):
        state, pre = env.reset(seed=seed, scramble_count=depth), None
        actions = []
        for t in range(1, MAX_T + 1):
            with torch.no_grad():
                x = torch.tensor(swap_orientation_labels(state) if control else state).float()
                a = model.get_action(x, pre) if mask else model.get_action(x)
            if mask:
                pre = a
            actions.append(a)
            state, _, done, _ = env.step(a)
            if done:
                return True, t, actions
        return False, 0, actions

    out = {"depths": np.array(DEPTHS), "seeds": np.array(SEEDS), "max_timesteps": np.array(MAX_T),
           "epoch": np.array(int(ck["epoch"])), "hidden": np.array(hidden)}
    for k, v in ck["model_state_dict"].items():
        out["w:" + k] = v.detach().cpu().numpy().astype(np.float32)
    for mask in (False, True):
        tag = "mask" if mask else "plain"
        solved = np.zeros((len(DEPTHS), len(SEEDS)), dtype=bool)
        steps = np.zeros((len(DEPTHS), len(SEEDS)), dtype=np.int64)
        acts = np.full((len(DEPTHS), len(SEEDS), MAX_T), -1, dtype=np.int8)
        for i, d in enumerate(DEPTHS):
            for j, s in enumerate(SEEDS):
                ok, t, al = episode(s, d, mask, False)
                solved[i, j], steps[i, j] = ok, t
                acts[i, j, :len(al)] = al
            print("%s depth %2d: reference loop solves %.1f %%" % (tag, d, 100.0 * solved[i].mean()), flush=True)
        out["solved_" + tag], out["steps_" + tag], out["actions_" + tag] = solved, steps, acts
    control = np.zeros(len(DEPTHS))
    for i, d in enumerate(DEPTHS):
        control[i] = np.mean([episode(s, d, False, True)[0] for s in SEEDS[:100]])
        print("control (orientation labels 1 <-> 2) depth %2d: %.1f %%" % (d, 100.0 * control[i]), flush=True)
    out["control_solve_rate"] = control

    # the reference's own MCTS (mcts.py, test.py:139-147: up to numMCTSSim = 50 simulations of one tree) with the
    # checkpoint as its value / policy net -- BASELINE config 5's consumer
    import importlib
    import random
    mcts_mod = importlib.import_module("mcts")
    cfg = {"mcts": {"numMCTSSim": 50, "cpuct": 1.0, "virtual_loss_const": 150, "value_min": -10.0}, "test": {"cube_size": 2}}
    cases = [(s, d) for d in (5, 8, 12) for s in range(40)]
    m_solved, m_sims, m_acts = [], [], []
    for seed, depth in cases:
        state = env.reset(seed=seed, scramble_count=depth)
        random.seed(1000 + seed)
        tree = mcts_mod.MCTS(model, cfg)
        result, used = None, 50
        with torch.no_grad():
            for k in range(50):
                result = tree.train(state, env)
                if result is not None:
                    used = k + 1
                    break
        acts = list(result) if result is not None else []
        m_solved.append(result is not None)
        m_sims.append(used)
        m_acts.append(acts + [-1] * (51 - len(acts)))
    out["mcts_cases"], out["mcts_solved"] = np.array(cases), np.array(m_solved)
    out["mcts_n_sims"], out["mcts_actions"] = np.array(m_sims), np.array(m_acts, dtype=np.int8)
    for d in (5, 8, 12):
        sel = out["mcts_cases"][:, 1] == d
        print("reference MCTS (50 simulations) depth %2d: solves %.1f %%" % (d, 100.0 * out["mcts_solved"][sel].mean()), flush=True)
    dst = os.path.join(os.path.dirname(HERE), "tests", "golden", "pin222.npz")
    np.savez_compressed(dst, **out)
    print("wrote", dst, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
