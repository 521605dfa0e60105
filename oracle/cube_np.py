"""Batched NumPy oracle for the cube hot path (TEST INFRASTRUCTURE ONLY).

Vectorised over instances; every function states the reference lines it
follows.  All integer results are exact, so parity against the CUDA path is
byte equality.  See oracle/__init__.py for who may import this.
"""
import numpy as np

from . import tables as T


def _check_size(size):
    if size not in (2, 3):
        raise NotImplementedError("cube_size %r" % (size,))   # cube_env.py:43-44


def solved_states(size, n):
    """initState (py222) / initState_3 (py333.py:211-218), n copies, uint8."""
    _check_size(size)
    return np.tile(T.SOLVED[size].astype(np.uint8), (n, 1))


def apply_moves(size, states, actions):
    """One transition per instance: s[moveDefs[a]] (py333.py:220-222; py222 doMove).

    states [N,S] uint8, actions [N] ints in [0,A).  Returns a fresh [N,S] array.
    """
    _check_size(size)
    actions = np.asarray(actions)
    if actions.size and (actions.min() < 0 or actions.max() >= T.N_ACTIONS[size]):
        raise IndexError("action out of range")              # cube_env.py:86,96 (list index)
    rows = T.MOVE_DEFS[size][actions.astype(np.int64)]        # [N,S]
    return np.take_along_axis(np.asarray(states), rows, axis=1)


def is_solved(size, states):
    """Face uniformity (py333.py:229-233 isSolved_3; py222 isSolved)."""
    _check_size(size)
    states = np.asarray(states)
    k = T.N_STICKERS[size] // 6
    faces = states.reshape(states.shape[0], 6, k)
    return (faces == faces[:, :, :1]).all(axis=(1, 2))


def rewards(solved):
    """+1.0 when solved else -1.0 (cube_env.py:89-94, 99-104)."""
    return np.where(solved, np.float32(1.0), np.float32(-1.0)).astype(np.float32)


def scramble(size, moves, init=None, per_step=False):
    """Apply moves[:,0], moves[:,1], ... in order (cube_env.py:65-67, 189-191).

    moves [N,depth]; init [N,S] or None (= solved).  Returns final states, or
    (final, all_states[N,depth,S], solved[N,depth]) when per_step.
    """
    _check_size(size)
    moves = np.asarray(moves)
    n, depth = moves.shape
    s = solved_states(size, n) if init is None else np.array(init, dtype=np.uint8, copy=True)
    if not per_step:
        for k in range(depth):
            s = apply_moves(size, s, moves[:, k])
        return s
    trail = np.empty((n, depth, T.N_STICKERS[size]), dtype=np.uint8)
    flags = np.empty((n, depth), dtype=bool)
    for k in range(depth):
        s = apply_moves(size, s, moves[:, k])
        trail[:, k] = s
        flags[:, k] = is_solved(size, s)
    return s, trail, flags


def get_op(size, states):
    """(piece, orientation) per slot.

    3x3x3: getOP_3 (py333.py:224-227) -> [N,20,2], corners then edges, table
    holes left at (0,0).  2x2x2: py222 getOP -> [N,7,2].
    """
    _check_size(size)
    s = np.asarray(states).astype(np.int64)
    if size == 3:
        hc = s[:, T.CORNER_DEFS_3] @ T.CORNER_HASH_W          # [N,8]
        he = s[:, T.EDGE_DEFS_3] @ T.EDGE_HASH_W              # [N,12]
        return np.concatenate((T.CORNER_INDS_3[hc], T.EDGE_INDS_3[he]), axis=1)
    h = s[:, T.PIECE_DEFS_2] @ T.HASH_W_2
    return T.PIECE_INDS_2[h]


def onehot_columns(size, states):
    """Index of the single 1 in every row of the net input.

    3x3x3 (pos_to_state_3, py333.py:235-246): row = slot, column =
    piece*3+ori (corners) / piece*2+ori (edges) -> [N,20].
    2x2x2 (cube_env.py:142-147): row = cubelet id, column = position*3+ori -> [N,7].
    """
    op = get_op(size, states)
    if size == 3:
        cols = np.empty(op.shape[:2], dtype=np.int64)
        cols[:, :8] = op[:, :8, 0] * 3 + op[:, :8, 1]
        cols[:, 8:] = op[:, 8:, 0] * 2 + op[:, 8:, 1]
        return cols
    n = op.shape[0]
    cols = np.full((n, 7), -1, dtype=np.int64)
    pos = np.arange(7)[None, :] * 3 + op[:, :, 1]
    np.put_along_axis(cols, op[:, :, 0], pos, axis=1)
    return cols


def encode(size, states, dtype=np.uint8):
    """One-hot net input [N,R,C] (cube_env.py:132-152)."""
    r, c = T.STATE_DIM[size]
    states = np.asarray(states)
    n = states.shape[0]
    out = np.zeros((n, r, c), dtype=dtype)
    if size == 3:
        cols = onehot_columns(3, states)
        out[np.arange(n)[:, None], np.arange(r)[None, :], cols] = 1
    else:
        # written position by position so that a (non-reachable) state naming the
        # same cubelet twice sets two cells, exactly as the reference loop does
        op = get_op(2, states)
        for p in range(7):
            out[np.arange(n), op[:, p, 0], 3 * p + op[:, p, 1]] = 1
    return out


def decode_2(onehot):
    """state_to_sim_state for 2x2x2 (cube_env.py:154-175 + py222 getStickers)."""
    onehot = np.asarray(onehot)
    n = onehot.shape[0]
    col = onehot.argmax(axis=2)                               # first 1 per cubelet row
    out = np.zeros((n, 24), dtype=np.uint8)
    for pos, colour in T.FIXED_STICKERS_2:
        out[:, pos] = colour
    home = T.SOLVED[2][T.PIECE_DEFS_2]                        # [7,3]
    for cubelet in range(7):
        position, ori = col[:, cubelet] // 3, col[:, cubelet] % 3
        for k in range(3):
            # np.roll(home, ori)[k] == home[(k - ori) % 3]
            out[np.arange(n), T.PIECE_DEFS_2[position, k]] = home[cubelet, (k - ori) % 3]
    return out


def expand(size, states):
    """All A children of every state, in action order (cube_env.py:212-238).

    Returns children [N,A,S] u8, solved [N,A] bool.  (The reference stops at the
    first solved child; that override is applied by `adi_targets`.)
    """
    _check_size(size)
    states = np.asarray(states)
    a = T.N_ACTIONS[size]
    children = states[:, T.MOVE_DEFS[size]]                   # [N,A,S]
    n = states.shape[0]
    sol = is_solved(size, children.reshape(n * a, -1)).reshape(n, a)
    return children.astype(np.uint8), sol


def first_solved_child(solved):
    """Index of the first solved child per row, -1 if none (cube_env.py:217-220)."""
    solved = np.asarray(solved)
    any_ = solved.any(axis=1)
    return np.where(any_, solved.argmax(axis=1), -1)


def adi_targets(child_values, solved, parent_values, depth, temperature):
    """ADI target assembly (cube_env.py:239-252) from net outputs.

    child_values [N,A] = V(child); solved [N,A]; parent_values [N] = V(state);
    depth [N] scramble counts.  First solved child a -> (1.0, a); otherwise
    max_a(V(child_a) - 1.0), first max wins.
    """
    child_values = np.asarray(child_values, dtype=np.float32)
    value = child_values + np.float32(-1.0)
    tp = value.argmax(axis=1)
    tv = value[np.arange(value.shape[0]), tp]
    fs = first_solved_child(solved)
    tv = np.where(fs >= 0, np.float32(1.0), tv)
    tp = np.where(fs >= 0, fs, tp)
    weight = np.asarray(depth, dtype=np.float64) ** (-1.0 * temperature)
    err = np.abs(np.asarray(parent_values, dtype=np.float64) - tv) * weight
    return tv, tp, err


def reference_moves(size, seed, depth):
    """The move sequence `reset(seed, depth)` draws (cube_env.py:62-68)."""
    return np.random.RandomState(seed).randint(T.N_ACTIONS[size], size=depth)


# ---- opt-in EXACT 3x3x3 encoding (not in the reference: the spec of include/cube_b200.h CUBE_ENCODING_EXACT) ----
# The shipped corner table is lossy (row 6 of corner_pieceDefs is read the wrong way round, 21 reachable
# hashes unassigned, py333.py:140-180).  The exact encoding keeps the reference's layout and hash weights,
# reads slot 6 like the other seven slots and assigns all 24 rotations the py222 way:
# (piece, ori) <- np.roll(home colours of the piece, ori).  Derived here from this oracle's own constants,
# independently of the product's table generator.
def _exact_corner_tables():
    defs = T.CORNER_DEFS_3.copy()
    defs[6] = defs[6][[0, 2, 1]]                                # [29, 15, 26]
    home = T.SOLVED[3][defs]                                    # [8, 3] colours of every piece at home
    inds = np.full((62, 2), -1, dtype=np.int64)
    for p in range(8):
        for o in range(3):
            h = int(np.roll(home[p], o) @ T.CORNER_HASH_W)
            assert inds[h, 0] < 0, "the exact corner hash must be injective"
            inds[h] = (p, o)
    return defs, home, inds


CORNER_DEFS_EXACT, CORNER_HOME_EXACT, CORNER_INDS_EXACT = _exact_corner_tables()


def onehot_columns_exact(states):
    """[N, 20] column of the 1 per row in the EXACT 3x3x3 encoding; -1 where a corner triple is no cubie at all."""
    s = np.asarray(states).astype(np.int64)
    hc = s[:, CORNER_DEFS_EXACT] @ T.CORNER_HASH_W
    op = CORNER_INDS_EXACT[hc]
    cols = np.empty((s.shape[0], 20), dtype=np.int64)
    cols[:, :8] = np.where(op[:, :, 0] >= 0, op[:, :, 0] * 3 + op[:, :, 1], -1)
    he = s[:, T.EDGE_DEFS_3] @ T.EDGE_HASH_W
    eo = T.EDGE_INDS_3[he]
    cols[:, 8:] = eo[:, :, 0] * 2 + eo[:, :, 1]
    return cols


def encode_exact(states, dtype=np.uint8):
    cols = onehot_columns_exact(states)
    n = cols.shape[0]
    out = np.zeros((n, 20, 24), dtype=dtype)
    out[np.arange(n)[:, None], np.arange(20)[None, :], np.maximum(cols, 0)] = 1
    return out


def decode_exact(onehot):
    """Inverse of encode_exact: [N, 20, 24] -> sticker rows [N, 54]."""
    onehot = np.asarray(onehot)
    n = onehot.shape[0]
    col = onehot.argmax(axis=2)
    out = np.zeros((n, 54), dtype=np.uint8)
    out[:, 4::9] = np.arange(6, dtype=np.uint8)                 # centres never move
    rows = np.arange(n)
    for q in range(8):
        piece, ori = col[:, q] // 3, col[:, q] % 3
        for k in range(3):                                      # np.roll(home, ori)[k] == home[(k - ori) % 3]
            out[rows, CORNER_DEFS_EXACT[q, k]] = CORNER_HOME_EXACT[piece, (k - ori) % 3]
    edge_home = T.SOLVED[3][T.EDGE_DEFS_3]                      # [12, 2]
    for q in range(12):
        piece, ori = col[:, 8 + q] // 2, col[:, 8 + q] % 2
        for k in range(2):
            out[rows, T.EDGE_DEFS_3[q, k]] = edge_home[piece, (k - ori) % 2]
    return out
