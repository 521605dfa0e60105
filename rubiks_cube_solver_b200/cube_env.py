"""Drop-in `CubeEnv`: the reference's gym-cube environment
(gym-cube/gym_cube/envs/cube_env.py:12-275) with every cube computation done by
the CUDA kernels through the C ABI.

Same constructor, attributes, return types and RNG discipline as the reference,
so `train.py` (:141,155,186,191), `mcts.py` (:37,80,96-101) and `test.py`
(:39,123,138) run unchanged:

* ``step`` -> ``(obs, reward, done, {})``; obs is ``np.float64 [7,21]`` (2x2x2) or
  ``np.int64 [20,24]`` (3x3x3); reward is the Python float +-1.0.
* ``reset(seed=None, scramble_count=2)`` draws its moves from the legacy global
  NumPy RNG exactly as cube_env.py:61-68 does (state saved and restored around it)
  and returns only the observation; ``scramble_count=0`` raises UnboundLocalError
  like cube_env.py:69.
* ``sim_cube`` (int64 sticker row) is the single source of truth and lives on the
  host, so ``copy.deepcopy(env)`` (mcts.py:37,96,101) is cheap and callers that
  assign ``env.sim_cube`` keep working; the device only ever sees copies.

Rendering (`render`, `close_render`, `save_video`, cube_env.py:113-130, 254-275) is
the reference's matplotlib GUI and is out of scope: those methods raise
NotImplementedError.

This class moves ONE cube per call and is therefore latency-bound by design; the
throughput path is `BatchedCubeEnv` / `ops` / `adi`, which `get_random_samples`
below already uses internally.
"""
import threading

import numpy as np
import torch

from . import ops

try:                                    # classic gym is optional (not installed here)
    import gym as _gym
    _EnvBase = _gym.Env
except Exception:                       # noqa: BLE001
    _EnvBase = object

ACTION_NAMES = {
    2: ["U", "U'", "F", "F'", "R", "R'"],
    3: ["U", "U'", "F", "F'", "R", "R'", "D", "D'", "B", "B'", "L", "L'"],
    'render': [["U", 1], ["U", -1], ["F", 1], ["F", -1], ["R", 1], ["R", -1],
               ["D", 1], ["D", -1], ["B", 1], ["B", -1], ["L", 1], ["L", -1]],
}


def get_env_config(cube_size=3):
    """utils.py:162-186."""
    if cube_size == 2:
        return [7, 21], 6
    if cube_size == 3:
        return [20, 24], 12
    raise NotImplementedError


def _sim_device(device):
    """CUDA device the kernels run on.  The env's `device` attribute keeps the reference's
    meaning (where model tensors are created, cube_env.py:240,250) and may be the CPU."""
    if not torch.cuda.is_available():
        raise RuntimeError("rubiks_cube_solver_b200 needs a CUDA device: there is no CPU fallback")
    d = torch.device(device) if device is not None else torch.device("cuda")
    if d.type == "cuda":
        return torch.device("cuda", d.index if d.index is not None else torch.cuda.current_device())
    return torch.device("cuda", torch.cuda.current_device())


_private_rng = threading.local()          # one legacy generator per thread, reseeded by every reset(seed, ...)


class CubeEnv(_EnvBase):
    metadata = {'render_modes': ['human', 'rgb_array']}

    def __init__(self, device, cube_size=2):
        self.cube_size = cube_size
        self.device = device
        self.action_to_sim_action = {k: (list(v) if k != 'render' else [list(x) for x in v])
                                     for k, v in ACTION_NAMES.items()}
        self.show_cube = False
        self.state_dim, self.action_dim = get_env_config(cube_size)
        self._sim_device = _sim_device(device)
        self.init_state()

    # ------------------------------------------------------------------ device helpers
    def _upload(self, sim_state):
        arr = np.ascontiguousarray(sim_state, dtype=np.uint8).reshape(1, -1)
        return torch.from_numpy(arr).to(self._sim_device)

    def _host(self):
        # the calling thread's single-cube session (mapped pinned page, C ABI cube_env_host_*): shared
        # by all envs of this size and device, so a deepcopy of the env copies no handle
        return ops.host_cube(self.cube_size, self._sim_device.index)

    def _sim_u8(self, sim_state):
        arr = np.ascontiguousarray(sim_state, dtype=np.uint8).reshape(-1)
        if arr.shape[0] != ops.N_STICKERS[self.cube_size]:
            raise ValueError("a sticker array of %d entries is expected" % ops.N_STICKERS[self.cube_size])
        return arr

    def _obs_from_u8(self, onehot_u8):
        if self.cube_size == 2:
            return onehot_u8.astype(np.float64)               # np.zeros(state_dim), cube_env.py:141
        return onehot_u8.astype(np.int64)                     # np.zeros([20,24], dtype=np.int), py333.py:238

    # ------------------------------------------------------------------ reference interface
    def init_state(self):
        if self.cube_size not in (2, 3):
            raise NotImplementedError
        per = ops.N_STICKERS[self.cube_size] // 6
        self.sim_cube = np.repeat(np.arange(6), per)          # int64, like initState / initState_3
        self.cube = self.sim_state_to_state(self.sim_cube)

    def reset(self, seed=None, scramble_count=2):
        self.init_state()
        if seed is not None:
            # np.random.seed(seed); randint(...); set_state(origin) (cube_env.py:60-65) leaves the global generator
            # exactly as it was, and a private legacy generator seeded alike draws the same indices -- without
            # the get_state / seed / set_state round trip of the 624-word state (~100 us per call)
            rs = getattr(_private_rng, "rs", None)
            if rs is None:
                rs = _private_rng.rs = np.random.RandomState(0)
            rs.seed(seed)
            action_sequence = rs.randint(self.action_dim, size=scramble_count)
        else:
            origin_state = np.random.get_state()
            action_sequence = np.random.randint(self.action_dim, size=scramble_count)
            np.random.set_state(origin_state)
        if len(action_sequence) == 0:
            raise UnboundLocalError("local variable 'state' referenced before assignment")   # cube_env.py:69
        host = self._host()
        if len(action_sequence) <= host.max_depth:
            stickers, onehot, _ = host.scramble(np.ascontiguousarray(action_sequence, dtype=np.uint8))
        else:                                               # absurdly deep: through device tensors
            moves = torch.from_numpy(action_sequence.astype(np.uint8).reshape(1, -1)).to(self._sim_device)
            states, _, _ = ops.scramble(self.cube_size, moves)
            onehot = ops.encode(self.cube_size, states, dtype=torch.uint8)[0].cpu().numpy()
            stickers = states[0].cpu().numpy()
        self.sim_cube = stickers.astype(np.int64)
        self.cube = self._obs_from_u8(onehot)
        return self.cube

    def step(self, action):
        if self.cube_size not in (2, 3):
            raise NotImplementedError
        self.action_to_sim_action[self.cube_size][action]     # IndexError / TypeError like the reference's list lookup
        a = int(action) % self.action_dim
        stickers, onehot, done = self._host().step(self._sim_u8(self.sim_cube), a)
        self.sim_cube = stickers.astype(np.int64)
        self.cube = self._obs_from_u8(onehot)
        reward = 1.0 if done else -1.0
        if self.show_cube:
            raise NotImplementedError("rendering is out of scope (matplotlib GUI)")
        return self.cube, reward, done, {}

    def render(self, mode=None):
        raise NotImplementedError("rendering is the reference's matplotlib GUI and is out of scope")

    def close_render(self):
        raise NotImplementedError("rendering is the reference's matplotlib GUI and is out of scope")

    def save_video(self, cube_size, scramble_count, sample_cube_count, video_path='./video'):
        raise NotImplementedError("rendering is the reference's matplotlib GUI and is out of scope")

    def sim_state_to_state(self, sim_state):
        if self.cube_size not in (2, 3):
            raise NotImplementedError
        return self._obs_from_u8(self._host().encode(self._sim_u8(sim_state)))

    def state_to_sim_state(self, state):
        if self.cube_size == 2:
            onehot = torch.from_numpy(np.ascontiguousarray(np.asarray(state) == 1.0).astype(np.uint8)[None])
            stickers = ops.decode(2, onehot.to(self._sim_device))
            return stickers[0].cpu().numpy().astype(np.int64)     # np.zeros(24, int) in py222 getStickers
        raise NotImplementedError                             # cube_env.py:171-174

    def get_target_value(self, model, scramble_count, temperature):
        """cube_env.py:196-252 for the env's current cube."""
        res = ops.expand(self.cube_size, self._upload(self.sim_cube), dtype=torch.float32)
        solved = res["solved"][0].cpu().numpy().astype(bool)
        first = int(solved.argmax()) if solved.any() else -1
        if first >= 0:
            target_value, target_policy = 1.0, first
        else:
            next_state_tensor = res["child_onehot"][0].to(self.device)
            reward_tensor = torch.full((self.action_dim,), -1.0, device=self.device)
            with torch.no_grad():
                next_value, _ = model(next_state_tensor)
                value = next_value.squeeze(dim=-1).detach() + reward_tensor
            target_value, target_policy = torch.max(value, -1, keepdim=True)
            target_value, target_policy = target_value.item(), target_policy.item()
        weight = scramble_count ** (-1 * temperature)
        with torch.no_grad():
            state_tensor = torch.tensor(self.cube, device=self.device).float()
            value, _ = model(state_tensor)
            error = abs(value.detach().item() - target_value) * weight
        return target_value, target_policy, error

    def get_random_samples(self, replay_buffer, model, sample_scramble_count, sample_cube_count, temperature):
        """cube_env.py:177-194, batched: all cubes and all their scramble prefixes go through
        the GPU in one pass (adi.generate_samples); the dicts appended to `replay_buffer` have
        the reference's keys, order and host types."""
        from . import adi
        if sample_cube_count <= 0:
            return
        # same draws, in the same order, from the same (global, un-restored) RNG as the reference
        moves = np.stack([np.random.randint(self.action_dim, size=sample_scramble_count)
                          for _ in range(sample_cube_count)])
        if sample_scramble_count == 0:
            self.init_state()
            return
        batch = adi.generate_samples(self.cube_size, torch.from_numpy(moves.astype(np.uint8)).to(self._sim_device),
                                     model, temperature, model_device=self.device)
        # the parents' one-hot rows are 0 / 1 in the net's dtype: one byte per element crosses the bus
        obs = self._obs_from_u8(batch["state"].to(torch.uint8).cpu().numpy())
        tv = batch["target_value"].cpu().tolist()           # Python floats of the float32 values, like .item()
        tp = batch["target_policy"].cpu().tolist()
        err = batch["error"].cpu().tolist()
        depth = sample_scramble_count
        append = replay_buffer.append
        i = 0
        for c in range(sample_cube_count):
            for k in range(1, depth + 1):
                append({'state': obs[i], 'target_value': tv[i], 'target_policy': tp[i], 'scramble_count': k,
                        'error': err[i]})
                i += 1
        # the env is left on the last cube's final state, as the reference loop leaves it
        self.sim_cube = batch["final_stickers"][-1].cpu().numpy().astype(np.int64)
        self.cube = obs[-1]
