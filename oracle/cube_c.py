"""NumPy-facing wrapper over the plain-C oracle (TEST INFRASTRUCTURE ONLY)."""
import ctypes

import numpy as np

from . import build as _build
from . import tables as T


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def _u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def num_threads():
    return _build.load().orc_num_threads()


def set_threads(n):
    _build.load().orc_set_threads(int(n))


def scramble(size, moves, init=None, per_step=False):
    lib = _build.load()
    moves = _u8(moves)
    n, depth = moves.shape
    S = T.N_STICKERS[size]
    init_ = None if init is None else _u8(init)
    out = np.empty((n, S), dtype=np.uint8)
    solved = np.empty(n, dtype=np.uint8)
    reward = np.empty(n, dtype=np.float32)
    per = np.empty((n, depth), dtype=np.uint8) if per_step else None
    cnt = ctypes.c_int64(0)
    rc = lib.orc_scramble(size, _p(init_), _p(moves), n, depth, _p(out), _p(solved), _p(reward),
                          _p(per), ctypes.byref(cnt))
    if rc:
        raise IndexError("action out of range")
    res = (out, solved.astype(bool), reward, int(cnt.value))
    return res + (per.astype(bool),) if per_step else res


def step(size, states, actions):
    lib = _build.load()
    s = np.array(states, dtype=np.uint8, copy=True, order="C")
    a = _u8(actions)
    n = s.shape[0]
    solved = np.empty(n, dtype=np.uint8)
    reward = np.empty(n, dtype=np.float32)
    if lib.orc_step(size, _p(s), _p(a), n, _p(solved), _p(reward)):
        raise IndexError("action out of range")
    return s, solved.astype(bool), reward


def is_solved(size, states):
    lib = _build.load()
    s = _u8(states)
    solved = np.empty(s.shape[0], dtype=np.uint8)
    lib.orc_solved(size, _p(s), s.shape[0], _p(solved), None)
    return solved.astype(bool)


def onehot_columns(size, states):
    lib = _build.load()
    s = _u8(states)
    cols = np.empty((s.shape[0], T.STATE_DIM[size][0]), dtype=np.uint8)
    lib.orc_columns(size, _p(s), s.shape[0], _p(cols))
    return cols


def encode(size, states):
    lib = _build.load()
    s = _u8(states)
    r, c = T.STATE_DIM[size]
    out = np.empty((s.shape[0], r, c), dtype=np.uint8)
    lib.orc_encode_u8(size, _p(s), s.shape[0], _p(out))
    return out


def expand(size, states, want_children=True, want_cols=True):
    lib = _build.load()
    s = _u8(states)
    n, A, S, R = s.shape[0], T.N_ACTIONS[size], T.N_STICKERS[size], T.STATE_DIM[size][0]
    children = np.empty((n, A, S), dtype=np.uint8) if want_children else None
    cols = np.empty((n, A, R), dtype=np.uint8) if want_cols else None
    solved = np.empty((n, A), dtype=np.uint8)
    lib.orc_expand(size, _p(s), n, _p(children), _p(cols), _p(solved))
    return children, cols, solved.astype(bool)
