"""rubiks_cube_solver_b200 -- B200-native batched Rubik's-cube simulator.

A drop-in for the rollout / training-data hot path of SUNGBEOMCHOI/Rubiks-Cube-Solver:
`make_env` / `CubeEnv` keep the reference's gym-cube interface (env.py:3-6,
gym-cube/gym_cube/envs/cube_env.py), `BatchedCubeEnv`, `ops`, `adi`, `rollout` and
`mcts_batch` are the batched forms, and all cube arithmetic runs in hand-written sm_100a
CUDA kernels behind the C ABI of include/cube_b200.h.  There is no CPU fallback.
"""
from . import adi, dist, mcts_batch, ops, rollout   # noqa: F401
from ._lib import CubeLibraryError, build_library, load as load_library      # noqa: F401
from .batch_env import BatchedCubeEnv   # noqa: F401
from .cube_env import CubeEnv, get_env_config      # noqa: F401
from .env import make_env               # noqa: F401

__all__ = ["make_env", "CubeEnv", "BatchedCubeEnv", "get_env_config", "ops", "adi", "dist", "rollout",
           "mcts_batch", "load_library", "build_library", "CubeLibraryError"]
