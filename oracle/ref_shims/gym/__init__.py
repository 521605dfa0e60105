"""Stub of the classic-gym surface the reference touches (env.py:1,5;
gym_cube/__init__.py:1; cube_env.py:1,12).  Test-only."""
from .envs import registration as _registration


class Env(object):
    metadata = {}


def make(env_id, **kwargs):
    entry = _registration.registry[env_id]
    mod_name, cls_name = entry.split(":")
    import importlib
    mod = importlib.import_module(mod_name)
    return getattr(mod, cls_name)(**kwargs)
