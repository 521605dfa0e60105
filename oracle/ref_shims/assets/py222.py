"""Stand-in for the reference's missing ``assets/py222.py`` (imported at
cube_env.py:8; listed in gym_cube.egg-info/SOURCES.txt:15 but absent from the
tree).  It is MeepMoop/py222's API with string moves, restated from the
published algorithm on top of ``oracle.tables``.
Test-only: it lets the reference's own cube_env.py run for cube_size=2."""
import numpy as np

from oracle import tables as T

_IDX = {name: i for i, name in enumerate(T.ACTIONS[2])}


def initState():
    return T.SOLVED[2].copy()


def doMove(s, move):
    return s[T.MOVE_DEFS_2[_IDX[move]]]


def isSolved(s):
    for f in range(6):
        if not (s[4 * f:4 * f + 4] == s[4 * f]).all():
            return False
    return True


def getOP(s):
    return T.PIECE_INDS_2[np.dot(s[T.PIECE_DEFS_2], T.HASH_W_2)]


def getStickers(sOP):
    s = np.zeros(24, dtype=np.int64)
    for pos, colour in T.FIXED_STICKERS_2:
        s[pos] = colour
    for i in range(7):
        home = T.SOLVED[2][T.PIECE_DEFS_2[sOP[i, 0]]]
        s[T.PIECE_DEFS_2[i]] = np.roll(home, sOP[i, 1])
    return s


def printCube(s):
    print(s)
