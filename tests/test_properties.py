"""Property tests (hypothesis) of the oracle and of the CPU emulation of the kernels' per-thread
code: group properties of the face turns, expand[a] == step(a), encode/decode round trips,
and emulation == oracle on arbitrary move sequences (including the no-op index and ragged depths)."""
import ctypes
import os

import numpy as np
import pytest
from hypothesis import given, settings
from hypothesis import strategies as st

from oracle import cube_np as O
from oracle import tables as T

from test_host_emulation import LIB, emul  # noqa: F401  (module-scoped fixture builds the emulation library)

SIZES = st.sampled_from((2, 3))


def seqs(size, max_len=40):
    return st.lists(st.integers(0, T.N_ACTIONS[size] - 1), min_size=0, max_size=max_len)


@settings(max_examples=150, deadline=None)
@given(size=SIZES, data=st.data())
def test_inverse_sequence_restores_any_state(size, data):
    seq = data.draw(seqs(size))
    start = np.array(data.draw(st.lists(st.integers(0, 255), min_size=T.N_STICKERS[size], max_size=T.N_STICKERS[size])),
                     dtype=np.uint8)[None]
    there = O.scramble(size, np.array([seq], dtype=np.int64).reshape(1, len(seq)), init=start)
    back = O.scramble(size, (np.array([seq[::-1]], dtype=np.int64).reshape(1, len(seq)) ^ 1), init=there)
    assert (back == start).all()
    # a turn only permutes stickers: the multiset of bytes never changes
    assert sorted(there[0]) == sorted(start[0])


@settings(max_examples=100, deadline=None)
@given(size=SIZES, data=st.data())
def test_expand_equals_step_and_solved_iff_uniform(size, data):
    seq = data.draw(seqs(size, 25))
    s = O.scramble(size, np.array([seq], dtype=np.int64).reshape(1, len(seq)))
    children, sol = O.expand(size, s)
    for a in range(T.N_ACTIONS[size]):
        stepped = O.apply_moves(size, s, np.array([a]))
        assert (children[0, a] == stepped[0]).all()
        k = T.N_STICKERS[size] // 6
        uniform = all(len(set(stepped[0, f * k:(f + 1) * k])) == 1 for f in range(6))
        assert bool(sol[0, a]) == uniform
    # reachable solved states are THE solved state (fixed centres / fixed DBL cubie)
    if O.is_solved(size, s)[0]:
        assert (s[0] == T.SOLVED[size]).all()


@settings(max_examples=100, deadline=None)
@given(seq=seqs(2, 30))
def test_2x2_encode_decode_round_trip(seq):
    s = O.scramble(2, np.array([seq], dtype=np.int64).reshape(1, len(seq)))
    enc = O.encode(2, s)
    assert enc.sum() == 7 and (enc.sum(axis=2) == 1).all()
    assert (O.decode_2(enc) == s).all()


@settings(max_examples=100, deadline=None)
@given(seq=seqs(3, 30))
def test_3x3_onehot_has_one_column_per_slot(seq):
    s = O.scramble(3, np.array([seq], dtype=np.int64).reshape(1, len(seq)))
    enc = O.encode(3, s)
    assert enc.shape == (1, 20, 24) and (enc.sum(axis=2) == 1).all()
    # edges are encoded bijectively even though corners are lossy
    assert len(set(np.argmax(enc[0, 8:], axis=1) // 2)) == 12


@settings(max_examples=120, deadline=None)
@given(size=SIZES, data=st.data())
def test_emulated_kernel_code_equals_oracle(emul, size, data):  # noqa: F811
    depth = data.draw(st.integers(0, 70))
    n = data.draw(st.integers(1, 40))
    a = T.N_ACTIONS[size]
    flat = data.draw(st.lists(st.integers(0, a), min_size=n * depth, max_size=n * depth))   # `a` itself -> no-op
    raw = np.array(flat, dtype=np.uint8).reshape(n, depth)
    moves = np.where(raw == a, 12, raw).astype(np.uint8)
    out = np.empty((n, T.N_STICKERS[size]), dtype=np.uint8)
    solved = np.empty(n, dtype=np.uint8)
    emul.emul_scramble(size, moves.ctypes.data_as(ctypes.c_void_p), n, depth, out.ctypes.data_as(ctypes.c_void_p),
                       solved.ctypes.data_as(ctypes.c_void_p))
    want = O.solved_states(size, n)
    for k in range(depth):
        live = moves[:, k] != 12
        if live.any():
            want[live] = O.apply_moves(size, want[live], moves[live, k])
    assert (out == want).all() and (solved.astype(bool) == O.is_solved(size, want)).all()
