"""One-cube-at-a-time CPU env with the reference's semantics and cost model
(TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

This is the "port" CPU baseline that bench.py times on the GPU box's host
cores (the box has no /root/reference): like the reference it keeps one cube
per env object, applies a move as a NumPy fancy-index gather, re-encodes the
one-hot observation with small NumPy calls plus a Python loop per slot, and
checks the six faces in a Python loop -- the work `CubeEnv.step` does at
cube_env.py:71-111 through py333.py:220-246 / py222.  It is also the scalar
cross-check of oracle/cube_np.py in the CPU tests.
"""
import numpy as np

from . import tables as T


class ScalarCubeEnv(object):
    """Single-cube env: reset / step / get_target_value / get_random_samples."""

    def __init__(self, cube_size=2, device=None):
        if cube_size not in (2, 3):
            raise NotImplementedError                          # cube_env.py:43-44
        self.cube_size = cube_size
        self.device = device
        self.state_dim = list(T.STATE_DIM[cube_size])
        self.action_dim = T.N_ACTIONS[cube_size]
        self._moves = T.MOVE_DEFS[cube_size]
        self._faces = T.N_STICKERS[cube_size] // 6
        self.init_state()

    # -- simulator primitives ------------------------------------------------
    def _solved(self, s):
        k = self._faces                                        # py333.py:229-233
        for f in range(6):
            face = s[k * f:k * f + k]
            if not (face == face[0]).all():
                return False
        return True

    def _observe(self, s):
        if self.cube_size == 3:                                # py333.py:224-227, 235-246
            corner = T.CORNER_INDS_3[np.dot(s[T.CORNER_DEFS_3], T.CORNER_HASH_W)]
            edge = T.EDGE_INDS_3[np.dot(s[T.EDGE_DEFS_3], T.EDGE_HASH_W)]
            obs = np.zeros((20, 24), dtype=np.int64)
            for slot, (piece, ori) in enumerate(np.concatenate((corner, edge))):
                obs[slot][piece * (3 if slot < 8 else 2) + ori] = 1
            return obs
        obs = np.zeros((7, 21))                                # cube_env.py:141-147 (float64)
        op = T.PIECE_INDS_2[np.dot(s[T.PIECE_DEFS_2], T.HASH_W_2)]
        for position, (cubelet, ori) in enumerate(op):
            obs[cubelet][position * 3 + ori] = 1.0
        return obs

    # -- reference-facing interface -------------------------------------------
    def init_state(self):
        self.sim_cube = T.SOLVED[self.cube_size].copy()       # cube_env.py:33-48
        self.cube = self._observe(self.sim_cube)

    def step(self, action):
        if not 0 <= int(action) < self.action_dim:
            raise IndexError("action out of range")
        self.sim_cube = self.sim_cube[self._moves[action]]    # cube_env.py:87,97
        self.cube = self._observe(self.sim_cube)
        done = self._solved(self.sim_cube)
        return self.cube, (1.0 if done else -1.0), done, {}

    def reset(self, seed=None, scramble_count=2):
        self.init_state()                                      # cube_env.py:50-69
        saved = np.random.get_state()
        if seed is not None:
            np.random.seed(seed)
        for action in np.random.randint(self.action_dim, size=scramble_count):
            state, _, _, _ = self.step(action)
        np.random.set_state(saved)
        return state

    def children(self):
        """Stickers, observation and solved flag of every child (cube_env.py:212-238),
        without the early break so callers can see all A of them."""
        out = []
        for a in range(self.action_dim):
            child = self.sim_cube[self._moves[a]]
            out.append((child, self._observe(child), self._solved(child)))
        return out

    def get_target_value(self, model, scramble_count, temperature):
        import torch                                           # cube_env.py:196-252
        obs, target = [], None
        for a in range(self.action_dim):
            child = self.sim_cube[self._moves[a]]
            o = self._observe(child)
            if self._solved(child):
                target = (1.0, a)
                break
            obs.append(o)
        if target is None:
            with torch.no_grad():
                x = torch.tensor(np.array(obs), device=self.device).float()
                v, _ = model(x)
                value = v.squeeze(dim=-1) + torch.full((len(obs),), -1.0, device=self.device)
            tv, tp = torch.max(value, -1, keepdim=True)
            target = (tv.item(), tp.item())
        with torch.no_grad():
            v, _ = model(torch.tensor(self.cube, device=self.device).float())
        err = abs(v.item() - target[0]) * scramble_count ** (-1 * temperature)
        return target[0], target[1], err

    def get_random_samples(self, replay_buffer, model, sample_scramble_count,
                           sample_cube_count, temperature):
        for _ in range(sample_cube_count):                     # cube_env.py:177-194
            self.init_state()
            seq = np.random.randint(self.action_dim, size=sample_scramble_count)
            for k, action in enumerate(seq):
                state, _, _, _ = self.step(action)
                tv, tp, err = self.get_target_value(model, k + 1, temperature)
                replay_buffer.append({'state': state, 'target_value': tv, 'target_policy': tp,
                                      'scramble_count': k + 1, 'error': err})
