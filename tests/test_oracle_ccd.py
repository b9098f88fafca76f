"""Validates the oracle's GJK-based CCD (conservative advancement under linear motion): closed-form
times of impact, and the defining property on random pairs — the gap is ~0 at the reported time of impact
and positive before it (first contact), positive throughout for misses.  The gap itself comes from the
GJK distance, which tests/test_oracle_narrow.py pins against closed forms and a QP."""
import numpy as np

import axcd
import oracle_lib as O


def _sweep(xa, sa, da, xb, sb, db, hull=None):
    xf = np.stack([xa, xb])
    sh = np.array([sa, sb], dtype=O.SHAPE_DT)
    return O.ccd_pairs(xf, sh, [[0, 1]], np.array([da, db], np.float32), hull)[0]


def _gap(xa, sa, xb, sb, hull=None):
    _, _, d, _ = O.collide_pair(xa, sa, xb, sb, hull)
    return d


def test_sphere_sphere_closed_form():
    xa, xb = O.xf((0, 0, 0)), O.xf((5, 0, 0))
    D = np.array([-6.0, 0.3, 0.0])
    r = _sweep(xa, O.sphere(0.5), (0, 0, 0), xb, O.sphere(0.7), D)
    c = np.array([5.0, 0, 0])
    a, b, cc = D @ D, 2 * c @ D, c @ c - 1.2 ** 2
    t = (-b - np.sqrt(b * b - 4 * a * cc)) / (2 * a)
    assert r["hit"] == 1 and abs(r["toi"] - t) < 1e-4
    n = (c + t * D) / np.linalg.norm(c + t * D)
    np.testing.assert_allclose([r["nx"], r["ny"], r["nz"]], n, atol=1e-3)
    # relative motion only: moving both by the same displacement changes nothing
    r2 = _sweep(xa, O.sphere(0.5), (1, 2, 3), xb, O.sphere(0.7), D + (1, 2, 3))
    assert r2["hit"] == 1 and abs(r2["toi"] - t) < 1e-4
    # a miss: passes beside
    r3 = _sweep(xa, O.sphere(0.5), (0, 0, 0), xb, O.sphere(0.7), (-6, 4.0, 0))
    assert r3["hit"] == 0 and r3["toi"] == 1.0
    # moving apart
    r4 = _sweep(xa, O.sphere(0.5), (0, 0, 0), xb, O.sphere(0.7), (3, 0, 0))
    assert r4["hit"] == 0
    # does not reach within the step
    r5 = _sweep(xa, O.sphere(0.5), (0, 0, 0), xb, O.sphere(0.7), (-3, 0, 0))
    assert r5["hit"] == 0
    # already overlapping
    r6 = _sweep(xa, O.sphere(0.5), (0, 0, 0), O.xf((0.9, 0, 0)), O.sphere(0.7), (-1, 0, 0))
    assert r6["hit"] == 1 and r6["toi"] == 0.0


def test_box_box_face_on_closed_form():
    xa = O.xf((0, 0, 0))
    xb = O.xf((4, 0.2, -0.1))
    r = _sweep(xa, O.box(0.5, 1, 1), (0.5, 0, 0), xb, O.box(1.0, 0.5, 0.5), (-3.5, 0, 0))
    # gap 2.5 closes at relative speed 4
    assert r["hit"] == 1 and abs(r["toi"] - 2.5 / 4.0) < 1e-4
    np.testing.assert_allclose([r["nx"], r["ny"], r["nz"]], [1, 0, 0], atol=1e-4)


def test_random_pairs_first_contact_property():
    rng = np.random.default_rng(3)
    s = axcd.generate_scene(400, 9, 30.0, frac_box=0.35, frac_sphere=0.25)   # 40 % hulls, far apart
    k = np.where(s.shapes["type"] == 0)[0][::2]
    s.shapes["type"][k] = 2
    s.shapes["p0"][k] = 0.25
    s.shapes["p1"][k] = 0.8
    hits = misses = 0
    kinds = set()
    for _ in range(300):
        a, b = rng.choice(s.n, 2, replace=False)
        xa, xb = s.xf[a].copy(), s.xf[b].copy()
        xb[:3] = xa[:3] + rng.normal(size=3) * 2.5
        if _gap(xa, s.shapes[a], xb, s.shapes[b], s.hull) <= 1e-3:
            continue
        aim = (xa[:3] - xb[:3]) * rng.uniform(0.6, 1.6) + rng.normal(size=3) * 0.5
        da = rng.normal(size=3).astype(np.float32) * 0.3
        db = (aim + da).astype(np.float32)
        r = _sweep(xa, s.shapes[a], da, xb, s.shapes[b], db, s.hull)

        def gap_at(t):
            pa, pb = xa.copy(), xb.copy()
            pa[:3] += np.float32(t) * da
            pb[:3] += np.float32(t) * db
            return _gap(pa, s.shapes[a], pb, s.shapes[b], s.hull)

        if r["hit"]:
            hits += 1
            kinds.add((int(s.shapes["type"][a]), int(s.shapes["type"][b])))
            assert -5e-4 < gap_at(r["toi"]) < 5e-4, (r, gap_at(r["toi"]))
            for t in np.linspace(0, r["toi"], 12)[:-1]:
                assert gap_at(t) > 0, (r, t)
        else:
            misses += 1
            assert r["toi"] == 1.0
            for t in np.linspace(0, 1, 40):
                assert gap_at(t) > 0, (r, t)
    assert hits > 60 and misses > 30, (hits, misses)
    assert len(kinds) >= 6


# ------------------------------------------------------------------ CCD with rotation -----------------
def _sweep_ang(xa, sa, da, wa, xb, sb, db, wb, hull=None):
    xf = np.stack([xa, xb])
    sh = np.array([sa, sb], dtype=O.SHAPE_DT)
    return O.ccd_pairs(xf, sh, [[0, 1]], np.array([da, db], np.float32), hull, rot=np.array([wa, wb], np.float32))[0]


def _gap_ang(xa, sa, da, wa, xb, sb, db, wb, t, hull=None):
    return _gap(O.ccd_pose_at(xa, da, wa, t), sa, O.ccd_pose_at(xb, db, wb, t), sb, hull)


def test_motion_model_is_first_order_quaternion_integration():
    q = O.axis_angle((0.3, -0.5, 0.8), 0.7)
    x = O.xf((1, 2, 3), q, (1.1, 0.9, 1.3))
    w = np.array([0.4, -0.2, 0.9], np.float32)
    for t in (0.0, 0.35, 1.0):
        got = O.ccd_pose_at(x, (0.5, -1.0, 2.0), w, t).astype(np.float64)
        qq = q.astype(np.float64)
        wq = np.r_[np.cross(w, qq[:3]) + qq[3] * w, -(w @ qq[:3])]      # (w, 0) (x) q, xyzw
        e = qq + 0.5 * t * wq
        np.testing.assert_allclose(got[3:7], e / np.linalg.norm(e), atol=2e-7)
        np.testing.assert_allclose(got[:3], np.array([1, 2, 3]) + t * np.array([0.5, -1.0, 2.0]), atol=1e-6)
        np.testing.assert_array_equal(got[7:], x[7:])
    # the turning angle reached at t = 1 is 2 atan(|w| / 2) <= |w|: the advancement's speed bound holds
    e1 = O.ccd_pose_at(x, (0, 0, 0), w, 1.0)[3:7].astype(np.float64)
    ang = 2 * np.arccos(min(1.0, abs(e1 @ q.astype(np.float64))))
    assert abs(ang - 2 * np.arctan(np.linalg.norm(w) / 2)) < 1e-5 and ang <= np.linalg.norm(w)


def test_spinning_bar_hits_a_resting_sphere():
    # a 2 x 0.2 x 0.2 bar turning about z sweeps its end into a sphere that linear CCD (no displacement) never sees
    bar, ball = O.box(1.0, 0.1, 0.1), O.sphere(0.25)
    xa, xb = O.xf((0, 0, 0)), O.xf((0.6, 0.75, 0))
    zero = (0, 0, 0)
    w = (0, 0, 1.4)
    assert _sweep(xa, bar, zero, xb, ball, zero)["hit"] == 0
    r = _sweep_ang(xa, bar, zero, w, xb, ball, zero, zero)
    assert r["hit"] == 1 and 0.05 < r["toi"] < 0.95 and r["iterations"] < 64
    assert -5e-4 < _gap_ang(xa, bar, zero, w, xb, ball, zero, zero, r["toi"]) < 5e-4
    ts = np.linspace(0, 1, 2001)
    gaps = np.array([_gap_ang(xa, bar, zero, w, xb, ball, zero, zero, t) for t in ts])
    first = ts[np.argmax(gaps <= 0)]
    assert gaps.min() < 0 and abs(first - r["toi"]) < 2e-3                # the first sampled overlap is at the reported time
    # turning the other way: the end moves away, no hit within the step
    r2 = _sweep_ang(xa, bar, zero, (0, 0, -1.4), xb, ball, zero, zero)
    assert r2["hit"] == 0 and r2["toi"] == 1.0
    # a sphere's own rotation changes nothing
    r3 = _sweep_ang(xa, bar, zero, w, xb, ball, zero, (3.0, -2.0, 1.0))
    assert r3["hit"] == 1 and abs(r3["toi"] - r["toi"]) < 1e-6


def test_zero_rotation_agrees_with_the_linear_sweep():
    rng = np.random.default_rng(5)
    s = axcd.generate_scene(200, 4, 30.0, frac_box=0.4, frac_sphere=0.2)
    pairs = rng.choice(s.n, (150, 2))
    pairs = pairs[pairs[:, 0] != pairs[:, 1]]
    xf = s.xf.copy()
    xf[pairs[:, 1], :3] = xf[pairs[:, 0], :3] + rng.normal(size=(len(pairs), 3)).astype(np.float32) * 2.5
    disp = rng.normal(size=(s.n, 3)).astype(np.float32) * 0.2
    disp[pairs[:, 1]] += (xf[pairs[:, 0], :3] - xf[pairs[:, 1], :3]) * rng.uniform(0.5, 1.5, (len(pairs), 1)).astype(np.float32)
    lin = O.ccd_pairs(xf, s.shapes, pairs, disp, s.hull)
    ang = O.ccd_pairs(xf, s.shapes, pairs, disp, s.hull, rot=np.zeros((s.n, 3), np.float32))
    assert np.array_equal(lin["hit"], ang["hit"]) and 20 < lin["hit"].sum() < len(pairs) - 10
    np.testing.assert_allclose(ang["toi"], lin["toi"], atol=2e-4)         # q is re-normalised: ULP-level pose differences


def test_random_turning_pairs_first_contact_property():
    rng = np.random.default_rng(13)
    s = axcd.generate_scene(400, 9, 30.0, frac_box=0.35, frac_sphere=0.25)
    k = np.where(s.shapes["type"] == 0)[0][::2]
    s.shapes["type"][k] = 2
    s.shapes["p0"][k] = 0.25
    s.shapes["p1"][k] = 0.8
    hits = misses = capped = 0
    kinds = set()
    for _ in range(260):
        a, b = rng.choice(s.n, 2, replace=False)
        xa, xb = s.xf[a].copy(), s.xf[b].copy()
        xb[:3] = xa[:3] + rng.normal(size=3) * 2.0
        sa, sb = s.shapes[a], s.shapes[b]
        if _gap(xa, sa, xb, sb, s.hull) <= 1e-3:
            continue
        aim = (xa[:3] - xb[:3]) * rng.uniform(0.3, 1.4) + rng.normal(size=3) * 0.4
        da = rng.normal(size=3).astype(np.float32) * 0.3
        db = (aim + da).astype(np.float32)
        wa, wb = (rng.normal(size=(2, 3)) * rng.uniform(0.0, 1.2)).astype(np.float32)
        r = _sweep_ang(xa, sa, da, wa, xb, sb, db, wb, s.hull)
        gap_at = lambda t: _gap_ang(xa, sa, da, wa, xb, sb, db, wb, t, s.hull)
        if r["hit"]:
            hits += 1
            kinds.add((int(sa["type"]), int(sb["type"])))
            if r["iterations"] < 64:
                assert -5e-4 < gap_at(r["toi"]) < 5e-4, (r, gap_at(r["toi"]))
            else:
                capped += 1                                              # conservative early hit: still no overlap before it
                assert gap_at(r["toi"]) > -5e-4
            for t in np.linspace(0, r["toi"], 25)[:-1]:
                assert gap_at(t) > 0, (r, t)
        else:
            misses += 1
            assert r["toi"] == 1.0
            for t in np.linspace(0, 1, 60):
                assert gap_at(t) > 0, (r, t)
    assert hits > 60 and misses > 30 and capped < 0.1 * hits, (hits, misses, capped)
    assert len(kinds) >= 6
