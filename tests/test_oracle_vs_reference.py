"""Oracle vs the reference's own code, imported live from /root/reference under the
test-only shims (oracle/ref_harness.py).  Skipped where the tree is absent (GPU box)."""
import numpy as np
import pytest

from oracle import cube_np as O
from oracle import ref_harness
from oracle import tables as T

pytestmark = pytest.mark.skipif(not ref_harness.reference_available(),
                                reason="/root/reference not present")


@pytest.fixture(scope="module")
def R():
    return ref_harness.load_reference()


def test_tables_live(R):
    p = R.py333
    assert (p.moveDefs == T.MOVE_DEFS_3).all()
    assert (p.corner_pieceDefs == T.CORNER_DEFS_3).all() and (p.edge_pieceDefs == T.EDGE_DEFS_3).all()
    assert (p.corner_pieceInds == T.CORNER_INDS_3).all() and (p.edge_pieceInds == T.EDGE_INDS_3).all()
    assert (p.corner_hashOP == T.CORNER_HASH_W).all() and (p.edge_hashOP == T.EDGE_HASH_W).all()
    assert [p.moveInds[n] for n in T.ACTIONS[3]] == list(range(12))
    assert R.utils.get_env_config(2) == ([7, 21], 6) and R.utils.get_env_config(3) == ([20, 24], 12)


@pytest.mark.parametrize("size", (2, 3))
def test_random_walks_live(R, size):
    env = R.make_env(size)
    rng = np.random.RandomState(99)
    A = T.N_ACTIONS[size]
    moves = rng.randint(A, size=(40, 25))
    _, trail, flags = O.scramble(size, moves, per_step=True)
    enc = O.encode(size, trail.reshape(40 * 25, -1)).reshape((40, 25) + T.STATE_DIM[size])
    for w in range(40):
        env.init_state()
        for k in range(25):
            obs, r, d, _ = env.step(int(moves[w, k]))
            assert (env.sim_cube == trail[w, k]).all()
            assert (obs == enc[w, k]).all()
            assert d == flags[w, k] and r == (1.0 if d else -1.0)


@pytest.mark.parametrize("size", (2, 3))
def test_reset_live(R, size):
    env = R.make_env(size)
    for seed, depth in ((0, 1), (3, 7), (50, 30), (1023, 10)):
        obs = env.reset(seed=seed, scramble_count=depth)
        final = O.scramble(size, O.reference_moves(size, seed, depth)[None, :])
        assert (env.sim_cube == final[0]).all()
        assert (obs == O.encode(size, final)[0]).all()
    with pytest.raises(UnboundLocalError):
        env.reset(seed=0, scramble_count=0)                  # cube_env.py:69
