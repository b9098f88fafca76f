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
