"""`BatchedCubeEnv`: N cubes stepped in lock-step on one GPU -- the throughput form of the
reference's `CubeEnv` (gym-cube/gym_cube/envs/cube_env.py).  Same action order, reward and
done semantics; observations are the one-hot network input written by the kernel in the
dtype the net consumes (bf16 by default)."""
import numpy as np
import torch

from . import ops


class BatchedCubeEnv(object):
    def __init__(self, n, cube_size=3, device=None, obs_dtype=torch.bfloat16):
        if cube_size not in (2, 3):
            raise NotImplementedError
        if not torch.cuda.is_available():
            raise RuntimeError("rubiks_cube_solver_b200 needs a CUDA device: there is no CPU fallback")
        self.n, self.cube_size, self.obs_dtype = int(n), cube_size, obs_dtype
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.state_dim, self.action_dim = list(ops.STATE_DIM[cube_size]), ops.N_ACTIONS[cube_size]
        self.counters = ops.new_counters(self.device)
        self.init_state()

    def init_state(self):
        self.sim_cube = ops.solved_states(self.cube_size, self.n, self.device)     # [N, S] uint8

    def observe(self, out=None):
        return ops.encode(self.cube_size, self.sim_cube, dtype=self.obs_dtype, out=out)

    @staticmethod
    def reference_moves(cube_size, seeds, scramble_count):
        """Host-side move draw identical to reset(seed, k) of the reference (cube_env.py:61-68)."""
        a = ops.N_ACTIONS[cube_size]
        return np.stack([np.random.RandomState(int(s)).randint(a, size=scramble_count) for s in seeds]).astype(np.uint8)

    def reset(self, seeds=None, scramble_count=2, moves=None, generator=None):
        """Scramble every cube from solved.  `moves` [N, k] (uint8, CUDA or host) wins; else
        `seeds` reproduces the reference's per-seed sequences; else i.i.d. uniform moves are
        drawn on the device (same distribution as np.random.randint, cube_env.py:65,189)."""
        if moves is None:
            if seeds is not None:
                if 0 < scramble_count <= 128:       # drawn on the device, bit-equal to RandomState(seed).randint
                    moves = ops.moves_from_seeds(self.cube_size, seeds, scramble_count, device=self.device)
                else:
                    moves = torch.from_numpy(self.reference_moves(self.cube_size, seeds, scramble_count))
            else:
                moves = torch.randint(0, self.action_dim, (self.n, scramble_count), dtype=torch.uint8,
                                      device=self.device, generator=generator)
        if not isinstance(moves, torch.Tensor):
            moves = torch.from_numpy(np.ascontiguousarray(moves, dtype=np.uint8))
        moves = moves.to(self.device).contiguous()
        if moves.shape[0] != self.n:
            raise ValueError("moves must have one row per cube")
        self.sim_cube, solved, reward = ops.scramble(self.cube_size, moves, counters=self.counters)
        return self.observe(), reward, solved

    def step(self, actions, validate=False):
        """actions [N] uint8.  Returns (obs, reward [N] float32, done [N] uint8, {})."""
        if not isinstance(actions, torch.Tensor):
            actions = torch.from_numpy(np.ascontiguousarray(actions, dtype=np.uint8))
        actions = actions.to(self.device, dtype=torch.uint8).contiguous()
        if validate:
            ops.validate_actions(self.cube_size, actions)
        _, solved, reward = ops.step(self.cube_size, self.sim_cube, actions, counters=self.counters)
        return self.observe(), reward, solved, {}

    def expand(self, **kw):
        return ops.expand(self.cube_size, self.sim_cube, dtype=self.obs_dtype, **kw)
