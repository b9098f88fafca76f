"""capsules and collision filtering.

Capsule placement follows the reference's drawCapsule (src/debug/physics_debug_draw.cpp:254-266:
local Y axis, endpoints (0, -+height/2, 0) through transformPoint, unscaled radius); filtering
follows gui::FilterInfo (include/axiom/gui/body_inspector.hpp:38-42)."""
import numpy as np
import pytest

import axcd
import oracle_lib as O
from test_oracle_narrow import qp_distance, box_vertices, TOL


def capsule_ends(t, height):
    h = np.float32(height) * np.float32(0.5)
    return (O.transform_point(t, (0, -h, 0)).astype(float), O.transform_point(t, (0, h, 0)).astype(float))


def mixed_scene(n=3000, seed=21, L=14.0):
    """boxes / spheres / capsules / hulls, roughly a quarter each."""
    s = axcd.generate_scene(n, seed, L, frac_box=0.5, frac_sphere=0.25)
    rng = np.random.default_rng(seed)
    box_idx = np.nonzero(s.shapes["type"] == 1)[0]
    cap = box_idx[rng.random(len(box_idx)) < 0.5]
    s.shapes["type"][cap] = 2
    s.shapes["p0"][cap] = rng.uniform(0.15, 0.35, len(cap)).astype(np.float32)   # radius
    s.shapes["p1"][cap] = rng.uniform(0.3, 1.0, len(cap)).astype(np.float32)     # height
    s.shapes["p2"][cap] = 0.0
    return s


# ------------------------------------------------------------------ oracle validation -----------
def test_capsule_refit_is_box_of_end_spheres():
    t = O.xf((1, 2, 3), O.axis_angle((0.2, -0.4, 0.9), 1.3), (1.5, 0.7, 2.0))
    rc, bb = O.refit([t], [O.capsule(0.25, 1.2)])
    a, b = capsule_ends(t, 1.2)
    lo = np.minimum(a, b) - 0.25
    hi = np.maximum(a, b) + 0.25
    np.testing.assert_allclose(bb[0], np.r_[lo, hi], atol=1e-6)


def test_capsule_vs_sphere_and_capsule_closed_forms():
    rng = np.random.default_rng(3)
    for _ in range(200):
        ta = O.xf(rng.uniform(0, 1.5, 3), O.axis_angle(rng.normal(size=3), rng.uniform(0, 6.28)))
        tb = O.xf(rng.uniform(0, 1.5, 3), O.axis_angle(rng.normal(size=3), rng.uniform(0, 6.28)))
        ra, ha, rb, hb = rng.uniform(0.1, 0.3), rng.uniform(0.2, 1.0), rng.uniform(0.1, 0.3), rng.uniform(0.2, 1.0)
        ea = np.array(capsule_ends(ta, ha))
        # capsule - sphere: point-to-segment distance
        hit, con, dist, _ = O.collide_pair(ta, O.capsule(ra, ha), tb, O.sphere(rb))
        d = qp_distance(ea, tb[:3].astype(float)[None, :]) - float(np.float32(ra)) - float(np.float32(rb))
        if abs(d) > 1e-3:
            assert hit == (d < 0)
            if qp_distance(ea, tb[:3].astype(float)[None, :]) > 1e-3:
                assert abs(dist - d) < TOL
        # capsule - capsule: segment-to-segment distance
        eb = np.array(capsule_ends(tb, hb))
        hit, con, dist, _ = O.collide_pair(ta, O.capsule(ra, ha), tb, O.capsule(rb, hb))
        core = qp_distance(ea, eb)
        d = core - float(np.float32(ra)) - float(np.float32(rb))
        if abs(d) > 1e-3 and core > 1e-3:
            assert hit == (d < 0)
            assert abs(dist - d) < TOL
            if hit:
                assert abs(con["depth"] + d) < TOL


def test_capsule_vs_box_against_qp():
    rng = np.random.default_rng(4)
    checked = 0
    for _ in range(150):
        ta = O.xf(rng.uniform(0, 1.6, 3), O.axis_angle(rng.normal(size=3), rng.uniform(0, 6.28)))
        tb = O.xf(rng.uniform(0, 1.6, 3), O.axis_angle(rng.normal(size=3), rng.uniform(0, 6.28)))
        r, h, hb = rng.uniform(0.1, 0.3), rng.uniform(0.2, 1.0), rng.uniform(0.25, 0.5, 3)
        hit, con, dist, epa = O.collide_pair(ta, O.capsule(r, h), tb, O.box(*hb))
        core = qp_distance(np.array(capsule_ends(ta, h)), box_vertices(tb[:3], tb[3:7], np.float32(hb).astype(float)))
        if core > 1e-3:
            d = core - float(np.float32(r))
            assert abs(dist - d) < TOL
            assert hit == (d <= 0) or abs(d) < TOL
            assert not epa
            checked += 1
        else:
            assert hit and epa and con["depth"] >= float(np.float32(r)) - TOL
    assert checked > 60


def test_filter_rule_semantics():
    bb = np.float32([[0, 0, 0, 1, 1, 1]] * 4)   # four coincident boxes: every pair overlaps
    allp = [[0, 1], [0, 2], [0, 3], [1, 2], [1, 3], [2, 3]]
    assert O.broadphase(bb, brute=True).tolist() == allp
    default = [(1, 0xFFFF, 0)] * 4                                   # FilterInfo defaults: everything collides
    assert O.broadphase(bb, brute=True, filters=default).tolist() == allp
    f = [(1, 2, 0), (2, 1, 0), (4, 0xFFFF, 0), (1, 0xFFFF, 0)]      # body 0 only sees category 2, body 1 only category 1
    assert O.broadphase(bb, brute=True, filters=f).tolist() == [[0, 1], [1, 3], [2, 3]]
    f = [(1, 0xFFFF, -3), (1, 0xFFFF, -3), (1, 0xFFFF, 5), (0, 0, 5)]   # same negative group never, same positive always
    assert O.broadphase(bb, brute=True, filters=f).tolist() == [[0, 2], [1, 2], [2, 3]]
    assert np.array_equal(O.broadphase(bb, filters=f), O.broadphase(bb, brute=True, filters=f))


def test_filtered_grid_equals_filtered_brute():
    s = mixed_scene(2000)
    rng = np.random.default_rng(9)
    filt = np.stack([1 << rng.integers(0, 4, s.n), rng.integers(1, 16, s.n), rng.integers(-2, 3, s.n)], axis=1)
    rc, bb = O.refit(s.xf, s.shapes, s.hull)
    assert rc == 0
    pf = O.broadphase(bb, brute=True, filters=filt)
    assert np.array_equal(pf, O.broadphase(bb, filters=filt, nthreads=3))
    assert 0 < len(pf) < len(O.broadphase(bb, brute=True))


# ------------------------------------------------------------------ GPU parity -------------------
@pytest.mark.gpu
def test_gpu_capsule_mix_bit_exact():
    from test_gpu_parity import run_and_compare
    s = mixed_scene(20000, seed=22, L=26.0)
    assert set(np.unique(s.shapes["type"])) == {0, 1, 2, 4}
    s.xf[:, 7:10] = np.random.default_rng(2).uniform(0.6, 1.4, (s.n, 3)).astype(np.float32)
    st, bitwise = run_and_compare(s)
    assert bitwise and st.numPenetrating > 0


@pytest.mark.gpu
def test_gpu_filters_match_oracle():
    s = mixed_scene(20000, seed=23, L=26.0)
    rng = np.random.default_rng(10)
    filt = np.stack([1 << rng.integers(0, 4, s.n), rng.integers(1, 16, s.n), rng.integers(-2, 3, s.n)], axis=1)
    w = axcd.CollisionWorld.for_scene(s)
    w.set_filters(filt)
    st = w.step()
    rc, bb = O.refit(s.xf, s.shapes, s.hull, nthreads=8)
    pairs = O.broadphase(bb, filters=filt, nthreads=8)
    con, _, _ = O.narrowphase(s.xf, s.shapes, pairs, s.hull, nthreads=8)
    assert np.array_equal(w.pairs(), pairs)
    gc = w.contacts()
    assert np.array_equal(gc, con)
    w.set_filters(None)                       # off again: the unfiltered set comes back
    st2 = w.step()
    assert st2.numPairs == len(O.broadphase(bb, nthreads=8)) > st.numPairs
    w.close()
