"""Oracle tables (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).

Restates the constant data of the reference simulators:

* 3x3x3 -- ``gym-cube/gym_cube/envs/assets/py333.py``: ``moveInds`` :41-44,
  ``moveDefs`` :46-138, ``corner_pieceDefs`` :140-149, ``edge_pieceDefs``
  :151-164, ``corner_hashOP``/``edge_hashOP`` :167-168, ``corner_pieceInds``
  :171-180, ``edge_pieceInds`` :182-198, ``initState_3`` :211-218.
* 2x2x2 -- the un-vendored ``assets/py222.py`` (MeepMoop/py222, unpinned;
  imported at ``cube_env.py:8``), restated from its published algorithm as in
  SURVEY.md Appendix A.
* action order -- ``cube_env.py:24-28``; ``get_env_config`` -- ``utils.py:162-186``.

The face turns are written here as sticker 4-cycles and expanded into gather
rows (``new[i] = old[row[i]]``, the semantics of ``s[moveDefs[m]]`` at
``py333.py:220-222``) instead of being listed as 12x54 literals;
``tests/test_oracle_vs_reference.py`` and the golden fixture
``tests/golden/reference_tables.npz`` pin the expansion against the reference's
own arrays.
"""
import numpy as np

# --------------------------------------------------------------------------
# action order (cube_env.py:24-28); even index = clockwise, odd = its inverse
# --------------------------------------------------------------------------
ACTIONS = {
    2: ["U", "U'", "F", "F'", "R", "R'"],
    3: ["U", "U'", "F", "F'", "R", "R'", "D", "D'", "B", "B'", "L", "L'"],
}
N_STICKERS = {2: 24, 3: 54}
N_ACTIONS = {2: 6, 3: 12}
STATE_DIM = {2: (7, 21), 3: (20, 24)}          # utils.py:162-186
ONEHOT_WIDTH = {2: 147, 3: 480}

# --------------------------------------------------------------------------
# 3x3x3 face turns as 4-cycles (a b c d): new[a]=old[b], new[b]=old[c], ...
# Sticker numbering: face f owns stickers 9f..9f+8, faces U,R,F,D,L,B
# (py333.py:3-19).  Five cycles per clockwise turn: two on the face itself,
# three on the ring of side stickers.
# --------------------------------------------------------------------------
_CW_CYCLES_3 = {
    "U": ((0, 6, 8, 2), (1, 3, 7, 5), (9, 45, 36, 18), (10, 46, 37, 19), (11, 47, 38, 20)),
    "F": ((6, 44, 29, 9), (7, 41, 28, 12), (8, 38, 27, 15), (18, 24, 26, 20), (19, 21, 25, 23)),
    "R": ((2, 20, 29, 51), (5, 23, 32, 48), (8, 26, 35, 45), (9, 15, 17, 11), (10, 12, 16, 14)),
    "D": ((15, 24, 42, 51), (16, 25, 43, 52), (17, 26, 44, 53), (27, 33, 35, 29), (28, 30, 34, 32)),
    "B": ((0, 11, 35, 42), (1, 14, 34, 39), (2, 17, 33, 36), (45, 51, 53, 47), (46, 48, 52, 50)),
    "L": ((0, 53, 27, 18), (3, 50, 30, 21), (6, 47, 33, 24), (36, 42, 44, 38), (37, 39, 43, 41)),
}


def _row_from_cycles(n, cycles):
    row = np.arange(n, dtype=np.int64)
    for cyc in cycles:
        for k, a in enumerate(cyc):
            row[a] = cyc[(k + 1) % len(cyc)]
    return row


def _invert(row):
    inv = np.empty_like(row)
    inv[row] = np.arange(len(row))
    return inv


def _build_move_defs_3():
    rows = []
    for name in ACTIONS[3][0::2]:
        cw = _row_from_cycles(54, _CW_CYCLES_3[name])
        rows.append(cw)
        rows.append(_invert(cw))
    return np.stack(rows)


MOVE_DEFS_3 = _build_move_defs_3()              # [12, 54]

# 2x2x2 sticker 4f+k sits where 3x3x3 sticker 9f+(0,2,6,8)[k] sits (same net,
# corners only); py222's U, U', F, F', R, R' rows are the 3x3x3 rows of the
# same name restricted to those stickers.
_CORNER_OF_2 = np.array([9 * (i // 4) + (0, 2, 6, 8)[i % 4] for i in range(24)])


def _build_move_defs_2():
    back = {int(s3): i for i, s3 in enumerate(_CORNER_OF_2)}
    rows = []
    for a in range(6):
        rows.append([back[int(MOVE_DEFS_3[a][s3])] for s3 in _CORNER_OF_2])
    return np.array(rows, dtype=np.int64)


MOVE_DEFS_2 = _build_move_defs_2()              # [6, 24]
MOVE_DEFS = {2: MOVE_DEFS_2, 3: MOVE_DEFS_3}

SOLVED = {                                       # initState / initState_3
    2: np.repeat(np.arange(6), 4).astype(np.int64),
    3: np.repeat(np.arange(6), 9).astype(np.int64),
}

# --------------------------------------------------------------------------
# 3x3x3 piece tables, exactly as shipped (py333.py:140-198).  Row 6 of the
# corner table lists its stickers in the opposite rotational sense to the other
# seven, and the 62-entry corner hash table leaves 38 entries at their
# np.zeros default, 21 of which reachable states hit: both quirks are part of
# the contract and are reproduced, not repaired.
# --------------------------------------------------------------------------
CORNER_DEFS_3 = np.array([
    [0, 47, 36], [6, 38, 18], [8, 20, 9], [2, 11, 45],
    [33, 42, 53], [27, 24, 44], [29, 26, 15], [35, 51, 17],
], dtype=np.int64)
EDGE_DEFS_3 = np.array([
    [1, 46], [3, 37], [7, 19], [5, 10], [34, 52], [30, 43],
    [28, 25], [32, 16], [21, 41], [23, 12], [48, 14], [50, 39],
], dtype=np.int64)
CORNER_HASH_W = np.array([1, 2, 10], dtype=np.int64)
EDGE_HASH_W = np.array([1, 10], dtype=np.int64)

# hash values the shipped table assigns to (piece p, orientation 0/1/2)
_CORNER_HASHES_3 = (
    (50, 54, 13), (28, 8, 42), (14, 5, 12), (52, 11, 15),
    (61, 44, 51), (47, 30, 40), (17, 35, 18), (23, 56, 21),
)


def _build_corner_inds_3():
    t = np.zeros((62, 2), dtype=np.int64)
    for p, hs in enumerate(_CORNER_HASHES_3):
        for o, h in enumerate(hs):
            t[h] = (p, o)
    return t


def _build_edge_inds_3():
    # edge p with home colours (a, b): a+10b -> (p,0), b+10a -> (p,1)
    t = np.zeros((55, 2), dtype=np.int64)
    for p, (s0, s1) in enumerate(EDGE_DEFS_3):
        a, b = s0 // 9, s1 // 9
        t[a + 10 * b] = (p, 0)
        t[b + 10 * a] = (p, 1)
    return t


CORNER_INDS_3 = _build_corner_inds_3()          # [62, 2]
EDGE_INDS_3 = _build_edge_inds_3()              # [55, 2]

# --------------------------------------------------------------------------
# 2x2x2 piece tables (py222 pieceDefs / hashOP / pieceInds, Appendix A)
# --------------------------------------------------------------------------
PIECE_DEFS_2 = np.array([
    [0, 21, 16], [2, 17, 8], [3, 9, 4], [1, 5, 20],
    [12, 10, 19], [13, 6, 11], [15, 22, 7],
], dtype=np.int64)
HASH_W_2 = np.array([1, 2, 10], dtype=np.int64)
FIXED_STICKERS_2 = ((14, 3), (18, 4), (23, 5))   # cubie DBL never moves under U/F/R


def _build_piece_inds_2():
    t = np.zeros((58, 2), dtype=np.int64)
    for p in range(7):
        c = SOLVED[2][PIECE_DEFS_2[p]]
        for o in range(3):
            t[int(np.dot(np.roll(c, o), HASH_W_2))] = (p, o)
    return t


PIECE_INDS_2 = _build_piece_inds_2()            # [58, 2]


def move_table(size):
    return MOVE_DEFS[size]
