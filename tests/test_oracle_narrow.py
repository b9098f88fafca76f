"""Independent validation of the oracle's AUTHORED GJK/EPA (the reference has none: parity unpinned).

Closed forms, a 15-axis SAT for box-box penetration depth and a scipy QP for
polytope distance.  Tolerance: 1e-4 absolute on lengths of O(1) shapes, the FP32 tolerance
BASELINE.json states.
"""
import numpy as np
import pytest
from scipy.optimize import minimize

import oracle_lib as O

TOL = 1e-4


def rot_matrix(q):
    return O.quat_to_mat3(q).astype(np.float64)


def box_vertices(pos, q, h):
    R = rot_matrix(q)
    s = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], float)
    return (s * np.asarray(h, float)) @ R.T + np.asarray(pos, float)


def qp_distance(VA, VB):
    """min |sum(a_i VA_i) - sum(b_j VB_j)| over the two simplices (convex QP, double precision)."""
    na, nb = len(VA), len(VB)
    M = np.vstack([VA, -VB])

    def f(x):
        d = x @ M
        return d @ d

    def g(x):
        return 2 * (M @ (x @ M))

    cons = [{"type": "eq", "fun": lambda x: x[:na].sum() - 1, "jac": lambda x: np.r_[np.ones(na), np.zeros(nb)]},
            {"type": "eq", "fun": lambda x: x[na:].sum() - 1, "jac": lambda x: np.r_[np.zeros(na), np.ones(nb)]}]
    x0 = np.r_[np.full(na, 1 / na), np.full(nb, 1 / nb)]
    best = None
    for _ in range(2):
        r = minimize(f, x0, jac=g, bounds=[(0, 1)] * (na + nb), constraints=cons, method="SLSQP",
                     options={"ftol": 1e-15, "maxiter": 500})
        x0 = r.x
        best = r.fun if best is None else min(best, r.fun)
    return float(np.sqrt(max(best, 0.0)))


def sat_box_box_depth(pa, qa, ha, pb, qb, hb):
    """Minimum translation distance of two overlapping boxes (15-axis SAT); <=0 when separated."""
    Ra, Rb = rot_matrix(qa), rot_matrix(qb)
    t = np.asarray(pb, float) - np.asarray(pa, float)
    axes = [Ra[:, i] for i in range(3)] + [Rb[:, i] for i in range(3)]
    for i in range(3):
        for j in range(3):
            c = np.cross(Ra[:, i], Rb[:, j])
            n = np.linalg.norm(c)
            if n > 1e-9:
                axes.append(c / n)
    best = np.inf
    for ax in axes:
        ra = sum(abs(ax @ Ra[:, i]) * ha[i] for i in range(3))
        rb = sum(abs(ax @ Rb[:, i]) * hb[i] for i in range(3))
        best = min(best, ra + rb - abs(ax @ t))
    return best


# ---------------------------------------------------------------- closed forms ----------------
def test_sphere_sphere_closed_form():
    rng = np.random.default_rng(1)
    for _ in range(200):
        c1, c2 = rng.uniform(-2, 2, 3), rng.uniform(-2, 2, 3)
        r1, r2 = rng.uniform(0.2, 1.5, 2)
        hit, c, dist, epa = O.collide_pair(O.xf(c1), O.sphere(r1), O.xf(c2), O.sphere(r2))
        c1f, c2f = np.float32(c1), np.float32(c2)
        d = np.linalg.norm((c2f - c1f).astype(np.float64))
        expect = d - np.float32(r1) - np.float32(r2)
        assert abs(dist - expect) < TOL
        assert hit == (dist <= 0)
        assert not epa
        if hit:
            n = (c2f - c1f) / d
            np.testing.assert_allclose([c["nx"], c["ny"], c["nz"]], n, atol=TOL)
            assert abs(c["depth"] + expect) < TOL
            mid = c1f + n * (np.float32(r1) + (d - np.float32(r1) - np.float32(r2)) / 2)
            np.testing.assert_allclose([c["px"], c["py"], c["pz"]], mid, atol=TOL)


def test_sphere_sphere_coincident_centres_fixed_fallback():
    hit, c, dist, _ = O.collide_pair(O.xf((1, 2, 3)), O.sphere(0.5), O.xf((1, 2, 3)), O.sphere(0.25))
    assert hit and c["depth"] == pytest.approx(0.75)
    assert (c["nx"], c["ny"], c["nz"]) == (1.0, 0.0, 0.0)   # vec_ops.hpp:202-208 safeNormalize fallback


def test_sphere_vs_axis_aligned_box_clamped_point():
    rng = np.random.default_rng(2)
    for _ in range(300):
        h = rng.uniform(0.3, 1.0, 3)
        c = rng.uniform(-2.5, 2.5, 3)
        r = rng.uniform(0.1, 0.8)
        hit, con, dist, epa = O.collide_pair(O.xf(), O.box(*h), O.xf(c), O.sphere(r))
        hf, cf, rf = np.float32(h).astype(float), np.float32(c).astype(float), float(np.float32(r))
        q = np.clip(cf, -hf, hf)
        outside = np.linalg.norm(cf - q)
        if outside > 1e-6:
            expect = outside - rf
            n = (cf - q) / outside
        else:   # centre inside the box: exit through the nearest face
            k = int(np.argmin(hf - np.abs(cf)))
            expect = -(hf[k] - abs(cf[k])) - rf
            n = np.zeros(3)
            n[k] = np.sign(cf[k]) if cf[k] != 0 else 1.0
        assert abs(dist - expect) < TOL, (h, c, r)
        assert hit == (expect <= 0) or abs(expect) < TOL
        if hit and abs(expect) > 1e-3 and (outside > 1e-3 or np.sort(hf - np.abs(cf))[1] - np.sort(hf - np.abs(cf))[0] > 1e-3):
            np.testing.assert_allclose([con["nx"], con["ny"], con["nz"]], n, atol=2e-4)
            assert abs(con["depth"] + expect) < TOL
        assert not epa   # sphere-box is closed form, inside or outside


def test_sphere_vs_rotated_scaled_box_closed_form():
    """Independent numpy restatement: clamp the sphere centre in the box frame (rotation from the
    quaternion, half lengths = halfExtent * scale), both argument orders."""
    rng = np.random.default_rng(12)
    for it in range(400):
        h = rng.uniform(0.2, 0.9, 3)
        sc = rng.uniform(0.5, 1.5, 3)
        axis = rng.normal(size=3)
        q = O.axis_angle(axis / np.linalg.norm(axis), rng.uniform(0, 2 * np.pi))
        cb = rng.uniform(-1, 1, 3)
        cs = cb + rng.uniform(-1.8, 1.8, 3)
        r = rng.uniform(0.1, 0.6)
        R = O.quat_to_mat3(q).astype(float)
        half = np.float32(h).astype(float) * np.float32(sc).astype(float)
        x = R.T @ (np.float32(cs).astype(float) - np.float32(cb).astype(float))
        k = np.clip(x, -half, half)
        out = np.linalg.norm(x - k)
        if np.any(np.abs(x) > half):
            expect = out - float(np.float32(r))
            n_box_to_sphere = R @ ((x - k) / out) if out > 0 else None
        else:
            gaps = half - np.abs(x)
            j = int(np.argmin(gaps))
            expect = -(gaps[j] + float(np.float32(r)))
            n_box_to_sphere = R[:, j] * (1.0 if x[j] >= 0 else -1.0)
            if np.sort(gaps)[1] - np.sort(gaps)[0] < 1e-3:
                n_box_to_sphere = None   # ambiguous exit face
        box_first = it % 2 == 0
        if box_first:
            hit, con, dist, epa = O.collide_pair(O.xf(cb, q, sc), O.box(*h), O.xf(cs), O.sphere(r))
        else:
            hit, con, dist, epa = O.collide_pair(O.xf(cs), O.sphere(r), O.xf(cb, q, sc), O.box(*h))
        assert not epa
        assert abs(dist - expect) < TOL
        if abs(expect) > 1e-3:
            assert hit == (expect < 0)
        if hit and abs(expect) > 1e-3:
            assert abs(con["depth"] + expect) < TOL
            if n_box_to_sphere is not None:
                n = n_box_to_sphere if box_first else -n_box_to_sphere   # normal points from A to B
                np.testing.assert_allclose([con["nx"], con["ny"], con["nz"]], n, atol=2e-4)
                # contact position = midpoint of the two surface witness points
                if out > 0 or not np.any(np.abs(x) > half):
                    surf_box = np.float32(cb).astype(float) + R @ (k if np.any(np.abs(x) > half) else x + (half[j] - abs(x[j])) * np.eye(3)[j] * (1 if x[j] >= 0 else -1))
                    surf_sph = np.float32(cs).astype(float) - n_box_to_sphere * float(np.float32(r))
                    np.testing.assert_allclose([con["px"], con["py"], con["pz"]], (surf_box + surf_sph) / 2, atol=5e-4)


# box-box runs through the closed-form SAT by default and through GJK/EPA with the generic flag
# (AXCD_FLAG_BOXBOX_GJK_EPA); both must give the closed-form answers below
BOXBOX_MODES = pytest.mark.parametrize("generic", [False, True], ids=["sat", "gjk_epa"])


@BOXBOX_MODES
def test_axis_aligned_box_box_min_overlap_axis(generic):
    rng = np.random.default_rng(3)
    for _ in range(300):
        ha, hb = rng.uniform(0.3, 1.0, 3), rng.uniform(0.3, 1.0, 3)
        t = rng.uniform(-2.0, 2.0, 3)
        hit, con, dist, epa = O.collide_pair(O.xf(), O.box(*ha), O.xf(t), O.box(*hb), cfg=O.default_cfg(True, generic))
        haf, hbf, tf = (np.float32(x).astype(float) for x in (ha, hb, t))
        gap = np.abs(tf) - haf - hbf
        if (gap > 0).any():
            expect = np.linalg.norm(np.maximum(gap, 0))
            assert not hit or expect < TOL
            assert abs(dist - expect) < TOL
        else:
            k = int(np.argmax(gap))
            assert hit and epa == generic
            assert abs(con["depth"] + gap[k]) < TOL
            srt = np.sort(gap)
            if srt[-1] - srt[-2] > 1e-3:
                n = np.zeros(3)
                n[k] = np.sign(tf[k])
                np.testing.assert_allclose([con["nx"], con["ny"], con["nz"]], n, atol=2e-4)


@BOXBOX_MODES
def test_box_rotated_45_touching_at_sqrt2(generic):
    # mirrors tests/math/aabb_test.cpp:330-349 (cube +-1 rotated 45 deg about Z reaches sqrt 2)
    q = O.axis_angle((0, 0, 1), np.pi / 4)
    r2 = float(np.sqrt(2.0))
    for gap in (0.25, 0.01, -0.01, -0.25):
        hit, con, dist, _ = O.collide_pair(O.xf(), O.box(1, 1, 1), O.xf((1 + r2 + gap, 0, 0), q), O.box(1, 1, 1),
                                           cfg=O.default_cfg(True, generic))
        assert abs(dist - gap) < TOL
        assert hit == (gap < 0)
        if hit:
            np.testing.assert_allclose([con["nx"], con["ny"], con["nz"]], (1, 0, 0), atol=2e-4)


@BOXBOX_MODES
def test_identical_shapes_identical_poses(generic):
    q = O.axis_angle((1, 2, 3), 0.7)
    hit, con, dist, epa = O.collide_pair(O.xf((1, 1, 1), q), O.box(0.5, 0.75, 1.0), O.xf((1, 1, 1), q), O.box(0.5, 0.75, 1.0),
                                         cfg=O.default_cfg(True, generic))
    assert hit and epa == generic and con["status"] == 0
    assert abs(con["depth"] - 1.0) < TOL     # thinnest direction: 2 * 0.5
    n = np.array([con["nx"], con["ny"], con["nz"]])
    assert abs(np.linalg.norm(n) - 1) < 1e-5
    ax = rot_matrix(q)[:, 0]
    assert abs(abs(n @ ax) - 1) < 1e-4


def test_point_like_hulls():
    hull = np.float32([[0, 0, 0], [0, 0, 0], [0, 0, 0], [0, 0, 0]])
    hit, con, dist, _ = O.collide_pair(O.xf((0, 0, 0)), O.hull_shape(0, 4), O.xf((0.5, 0, 0)), O.sphere(0.2), hull)
    assert not hit and abs(dist - 0.3) < TOL
    hit, con, dist, _ = O.collide_pair(O.xf((0, 0, 0)), O.hull_shape(0, 4), O.xf((0.1, 0, 0)), O.sphere(0.2), hull)
    assert hit and abs(con["depth"] - 0.1) < TOL
    hit, con, dist, _ = O.collide_pair(O.xf((0, 0, 0)), O.hull_shape(0, 4), O.xf((0, 0, 0)), O.hull_shape(0, 4), hull)
    assert hit and con["depth"] == 0.0


@BOXBOX_MODES
def test_stacked_boxes_resting_contact(generic):
    # exactly touching faces: either classification is float noise, but it must not blow up
    hit, con, dist, _ = O.collide_pair(O.xf((0, 0, 0)), O.box(1, 1, 1), O.xf((0.25, 2.0, -0.25)), O.box(1, 1, 1),
                                       cfg=O.default_cfg(True, generic))
    assert abs(dist) < TOL
    if hit:
        assert con["depth"] < TOL
        np.testing.assert_allclose([con["nx"], con["ny"], con["nz"]], (0, 1, 0), atol=1e-3)


# ---------------------------------------------------------------- randomized cross-checks -----
@BOXBOX_MODES
def test_random_box_box_vs_sat_and_qp(generic):
    rng = np.random.default_rng(4)
    n_pen = n_sep = 0
    for _ in range(400):
        ha, hb = rng.uniform(0.25, 0.5, 3), rng.uniform(0.25, 0.5, 3)
        pa, pb = rng.uniform(0, 1.5, 3), rng.uniform(0, 1.5, 3)
        qa = O.axis_angle(rng.normal(size=3), rng.uniform(0, 6.28))
        qb = O.axis_angle(rng.normal(size=3), rng.uniform(0, 6.28))
        haf, hbf = np.float32(ha).astype(float), np.float32(hb).astype(float)
        hit, con, dist, epa = O.collide_pair(O.xf(pa, qa), O.box(*ha), O.xf(pb, qb), O.box(*hb), cfg=O.default_cfg(True, generic))
        assert con["status"] == 0
        depth = sat_box_box_depth(np.float32(pa), qa, haf, np.float32(pb), qb, hbf)
        if depth > 0:
            n_pen += 1
            assert hit and epa == generic, (depth, dist)
            assert abs(con["depth"] - depth) < TOL
            # moving B by depth along n must separate (SAT depth ~ 0 afterwards)
            n = np.array([con["nx"], con["ny"], con["nz"]], float)
            assert abs(np.linalg.norm(n) - 1) < 1e-5
            after = sat_box_box_depth(np.float32(pa), qa, haf, np.float32(pb).astype(float) + n * (depth + 1e-3), qb, hbf)
            assert after < 1e-4
        else:
            n_sep += 1
            d = qp_distance(box_vertices(np.float32(pa), qa, haf), box_vertices(np.float32(pb), qb, hbf))
            assert not hit or d < TOL
            assert abs(dist - d) < TOL, (dist, d)
    assert n_pen > 30 and n_sep > 30


def test_random_hull_pairs_vs_qp():
    rng = np.random.default_rng(5)
    checked = 0
    for _ in range(120):
        va = (rng.normal(size=(16, 3)) * rng.uniform(0.25, 0.5, 3)).astype(np.float32)
        vb = (rng.normal(size=(16, 3)) * rng.uniform(0.25, 0.5, 3)).astype(np.float32)
        va /= np.maximum(1.0, np.linalg.norm(va, axis=1, keepdims=True) * 2)
        vb /= np.maximum(1.0, np.linalg.norm(vb, axis=1, keepdims=True) * 2)
        hull = np.vstack([va, vb])
        pa, pb = rng.uniform(0, 1.2, 3), rng.uniform(0, 1.2, 3)
        qa = O.axis_angle(rng.normal(size=3), rng.uniform(0, 6.28))
        qb = O.axis_angle(rng.normal(size=3), rng.uniform(0, 6.28))
        sa, sb = rng.uniform(0.5, 1.5, 3), rng.uniform(0.5, 1.5, 3)
        ta, tb = O.xf(pa, qa, sa), O.xf(pb, qb, sb)
        hit, con, dist, epa = O.collide_pair(ta, O.hull_shape(0, 16), tb, O.hull_shape(16, 16), hull)
        WA = np.array([O.transform_point(ta, v) for v in va], float)
        WB = np.array([O.transform_point(tb, v) for v in vb], float)
        d = qp_distance(WA, WB)
        if d > 1e-3:
            assert not hit
            assert abs(dist - d) < TOL
            checked += 1
        elif hit and con["depth"] > 1e-3:
            # separating B by depth along n must leave the hulls (nearly) touching
            n = np.array([con["nx"], con["ny"], con["nz"]], float)
            d2 = qp_distance(WA, WB + n * (con["depth"] + 2e-3))
            assert 1e-3 < d2 < 3e-3 + 1e-4, (d2, con["depth"])
            # and no shorter escape exists: shrinking the shift keeps them overlapping
            d3 = qp_distance(WA, WB + n * (con["depth"] - 2e-3))
            assert d3 < 1e-4
            checked += 1
    assert checked > 60


def test_sphere_hull_margin_contact_vs_qp():
    rng = np.random.default_rng(6)
    for _ in range(80):
        vb = (rng.normal(size=(12, 3)) * 0.3).astype(np.float32)
        pb = rng.uniform(0, 1.0, 3)
        qb = O.axis_angle(rng.normal(size=3), rng.uniform(0, 6.28))
        tb = O.xf(pb, qb)
        c, r = rng.uniform(0, 1.0, 3), rng.uniform(0.2, 0.5)
        hit, con, dist, epa = O.collide_pair(O.xf(c), O.sphere(r), tb, O.hull_shape(0, 12), vb)
        WB = np.array([O.transform_point(tb, v) for v in vb], float)
        d = qp_distance(np.float32(c).astype(float)[None, :], WB)
        if d > 1e-3:
            assert abs(dist - (d - float(np.float32(r)))) < TOL
            assert hit == (d - float(np.float32(r)) <= 0) or abs(d - r) < TOL
            assert not epa


def test_distances_mode_matches_default_contact_set():
    import axcd
    s = axcd.config_scene("C2", scale=0.003)
    rc, bb = O.refit(s.xf, s.shapes, s.hull)
    pairs = O.broadphase(bb)
    c1, _, st1 = O.narrowphase(s.xf, s.shapes, pairs, s.hull)
    c2, d2, st2 = O.narrowphase(s.xf, s.shapes, pairs, s.hull, want_distances=True)
    assert np.array_equal(c1, c2)
    assert st1.gjkIterations < st2.gjkIterations
    is_contact = np.zeros(len(pairs), bool)
    key = {(a, b) for a, b in zip(c2["a"], c2["b"])}
    for k, (a, b) in enumerate(pairs):
        is_contact[k] = (a, b) in key
    assert (d2[is_contact] <= 0).all() and (d2[~is_contact] > 0).all()


def test_epa_depth_is_the_global_minimum_vs_minkowski_hull():
    """Penetration depth = distance from the origin to the boundary of the Minkowski difference A - B.
    Independent restatement: scipy's convex hull of all pairwise vertex differences; the nearest facet
    plane gives the depth and the normal.  Pins EPA's optimality (not just validity along its own normal)."""
    from scipy.spatial import ConvexHull
    rng = np.random.default_rng(12)
    checked = 0
    for _ in range(150):
        va = (rng.normal(size=(12, 3)) * rng.uniform(0.3, 0.5, 3)).astype(np.float32)
        vb = (rng.normal(size=(12, 3)) * rng.uniform(0.3, 0.5, 3)).astype(np.float32)
        hull = np.vstack([va, vb])
        ta = O.xf(rng.uniform(0, 0.5, 3), O.axis_angle(rng.normal(size=3), rng.uniform(0, 6.28)), rng.uniform(0.7, 1.3, 3))
        tb = O.xf(rng.uniform(0, 0.5, 3), O.axis_angle(rng.normal(size=3), rng.uniform(0, 6.28)), rng.uniform(0.7, 1.3, 3))
        hit, con, dist, epa = O.collide_pair(ta, O.hull_shape(0, 12), tb, O.hull_shape(12, 12), hull)
        WA = np.array([O.transform_point(ta, v) for v in va], float)
        WB = np.array([O.transform_point(tb, v) for v in vb], float)
        md = ConvexHull((WA[:, None, :] - WB[None, :, :]).reshape(-1, 3))
        off = md.equations[:, 3]                      # n.x + off <= 0 inside; origin inside <=> all off <= 0
        if not (off < -1e-3).all():
            continue                                  # separated or grazing: covered by the QP tests
        k = int(np.argmax(off))
        depth = -off[k]
        assert hit and epa and con["status"] == 0
        assert abs(con["depth"] - depth) < 1e-3 * max(1.0, depth), (con["depth"], depth)
        # shifting B by depth along the facet's outward normal moves A - B the other way, so the origin leaves
        # through facet k: the contact normal (a to b) is that outward normal
        n = np.array([con["nx"], con["ny"], con["nz"]], float)
        if sorted(off)[-1] - sorted(off)[-2] > 1e-3:  # unique nearest facet: the normal is determined
            assert np.dot(n, md.equations[k, :3]) > 0.999, (n, md.equations[k, :3])
        checked += 1
    assert checked > 60, checked


def test_box_box_sat_vs_minkowski_hull_and_vs_epa():
    """The closed-form box-box answer (15-axis SAT) against two independent statements of the same quantity:
    the nearest facet of the convex hull of the Minkowski difference (scipy) — depth and normal are the
    global minimum — and the generic GJK/EPA path of this oracle on the same pair.  Also the witness points:
    each lies on its box's supporting plane for the contact normal."""
    from scipy.spatial import ConvexHull
    rng = np.random.default_rng(21)
    checked = 0
    for _ in range(300):
        ha, hb = rng.uniform(0.25, 0.5, 3), rng.uniform(0.25, 0.5, 3)
        sa, sb = rng.uniform(0.7, 1.4, 3), rng.uniform(0.7, 1.4, 3)
        qa = O.axis_angle(rng.normal(size=3), rng.uniform(0, 6.28))
        qb = O.axis_angle(rng.normal(size=3), rng.uniform(0, 6.28))
        ca = rng.uniform(0, 0.3, 3)
        cb = ca + rng.normal(size=3) * 0.35
        ta, tb = O.xf(ca, qa, sa), O.xf(cb, qb, sb)
        hit, con, dist, epa = O.collide_pair(ta, O.box(*ha), tb, O.box(*hb))
        hit2, con2, dist2, epa2 = O.collide_pair(ta, O.box(*ha), tb, O.box(*hb), cfg=O.default_cfg(True, True))
        assert not epa and (epa2 or not hit2)
        haf = np.float32(ha).astype(float) * np.float32(sa).astype(float)
        hbf = np.float32(hb).astype(float) * np.float32(sb).astype(float)
        WA = box_vertices(np.float32(ca), qa, haf)
        WB = box_vertices(np.float32(cb), qb, hbf)
        md = ConvexHull((WA[:, None, :] - WB[None, :, :]).reshape(-1, 3))
        off = md.equations[:, 3]
        if not (off < -1e-3).all():
            if (off > 1e-3).any():
                assert not hit and not hit2      # clearly apart
            continue
        k = int(np.argmax(off))
        depth = -off[k]
        assert hit and hit2
        assert abs(con["depth"] - depth) < 2e-5 * max(1.0, depth), (con["depth"], depth)   # the SAT is exact
        assert abs(con2["depth"] - depth) < 1e-3 * max(1.0, depth)                         # EPA within its tolerance
        n = np.array([con["nx"], con["ny"], con["nz"]], float)
        assert abs(np.linalg.norm(n) - 1) < 1e-5
        if sorted(off)[-1] - sorted(off)[-2] > 1e-3:
            assert np.dot(n, md.equations[k, :3]) > 0.9999
            assert np.dot(n, [con2["nx"], con2["ny"], con2["nz"]]) > 0.999
        # witness points: position = their midpoint, and they are depth apart along n on the supporting planes
        Ra, Rb = rot_matrix(qa), rot_matrix(qb)
        sup_a = float(np.sum(haf * np.abs(Ra.T @ n)))
        sup_b = float(np.sum(hbf * np.abs(Rb.T @ n)))
        pos = np.array([con["px"], con["py"], con["pz"]], float)
        # midpoint of a point on A's supporting plane (offset +sup_a along n from A's centre) and a point on
        # B's (offset -sup_b from B's centre): its n-coordinate is the mean of the two plane offsets
        expect = 0.5 * ((np.float32(ca).astype(float) @ n + sup_a) + (np.float32(cb).astype(float) @ n - sup_b))
        assert abs(pos @ n - expect) < 1e-4
        checked += 1
    assert checked > 100, checked
