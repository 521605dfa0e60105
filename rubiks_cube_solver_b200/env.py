"""`make_env(device, cube_size)` -- the reference's env factory (env.py:3-6).

Works without gym (not installed here).  When classic gym is importable, 'cube-v0' is
also registered with the reference's id and kwargs (gym_cube/__init__.py:4-7), but the
env is constructed directly so no wrapper can block `step` before `reset`
(cube_env.py:188-191 via train.py:155) or hide attributes such as `sim_cube`.
"""
from .cube_env import CubeEnv

ENV_ID = 'cube-v0'

try:
    from gym.envs.registration import register as _register
    try:
        _register(id=ENV_ID, entry_point='rubiks_cube_solver_b200.cube_env:CubeEnv')
    except Exception:       # noqa: BLE001 - already registered (e.g. by the reference's gym_cube)
        pass
except Exception:           # noqa: BLE001 - gym absent
    pass


def make_env(device, cube_size):
    return CubeEnv(device=device, cube_size=cube_size)
