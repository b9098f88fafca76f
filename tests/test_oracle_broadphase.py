"""The oracle's production (grid) broadphase against the brute-force truth."""
import numpy as np

import axcd
import oracle_lib as O


def test_grid_equals_brute_force_c0():
    s = axcd.config_scene("C0")
    rc, bb = O.refit(s.xf, s.shapes, s.hull)
    pb, pg = O.broadphase(bb, brute=True), O.broadphase(bb, nthreads=4)
    assert np.array_equal(pb, pg)
    assert len(pb) == 2875          # realised count for the C0 fixture (seed 1, N=1000, L=10)
    assert (pb[:, 0] < pb[:, 1]).all()
    key = pb[:, 0].astype(np.uint64) << np.uint64(32) | pb[:, 1]
    assert (np.diff(key.astype(np.int64)) > 0).all()


def test_grid_equals_brute_force_20k_hull_mix():
    s = axcd.config_scene("C2", scale=0.02)
    rc, bb = O.refit(s.xf, s.shapes, s.hull, nthreads=4)
    assert np.array_equal(O.broadphase(bb, brute=True), O.broadphase(bb, nthreads=4))


def test_grid_mixed_sizes_and_margin():
    rng = np.random.default_rng(0)
    n = 3000
    lo = rng.uniform(0, 20, (n, 3)).astype(np.float32)
    ext = (rng.uniform(0.01, 1.0, (n, 3)) ** 3 * 6).astype(np.float32)   # heavy-tailed sizes
    bb = np.hstack([lo, lo + ext])
    assert np.array_equal(O.broadphase(bb, brute=True), O.broadphase(bb, nthreads=3))


def test_touching_and_degenerate_boxes():
    bb = np.float32([[0, 0, 0, 1, 1, 1], [1, 0, 0, 2, 1, 1], [2.0000002, 0, 0, 3, 1, 1],
                     [0.5, 0.5, 0.5, 0.5, 0.5, 0.5], [np.nan, 0, 0, 1, 1, 1],
                     [0, 0, 0, np.inf, 1, 1]])
    pb = O.broadphase(bb, brute=True)
    pg = O.broadphase(bb)
    assert np.array_equal(pb, pg)
    assert pb.tolist() == [[0, 1], [0, 3], [0, 5], [1, 5], [2, 5], [3, 5]]


def test_worlds_never_pair_across():
    s = axcd.config_scene("C3", scale=8 / 4096)
    rc, bb = O.refit(s.xf, s.shapes, s.hull)
    pb = O.broadphase(bb, s.world_id, brute=True)
    pg = O.broadphase(bb, s.world_id)
    assert np.array_equal(pb, pg)
    assert (s.world_id[pb[:, 0]] == s.world_id[pb[:, 1]]).all()
    # and it equals the union of the per-world answers
    tot = 0
    for w in range(s.num_worlds):
        m = np.nonzero(s.world_id == w)[0]
        tot += len(O.broadphase(bb[m], brute=True))
    assert tot == len(pb)


def test_empty_and_single():
    assert len(O.broadphase(np.zeros((0, 6), np.float32))) == 0
    assert len(O.broadphase(np.float32([[0, 0, 0, 1, 1, 1]]))) == 0
