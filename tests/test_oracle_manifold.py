"""Validates the oracle's contact-manifold stage independently of its own code:
closed-form stacked boxes, an independent float64 numpy clipper (2-D rectangle clipping in the
reference face's own coordinates instead of 3-D plane clipping), and geometric properties on a
random scene.  The CUDA path is then compared with this oracle in tests/test_gpu_parity.py."""
import numpy as np
import pytest

import oracle_lib as O


def _contact(a, sa, b, sb):
    hit, c, _, _ = O.collide_pair(a, sa, b, sb)
    assert hit
    c = c.copy()
    c["a"], c["b"] = 0, 1
    return c


def _manifold(a, sa, b, sb):
    c = _contact(a, sa, b, sb)
    m, tot = O.manifolds(np.stack([a, b]), np.array([sa, sb], dtype=O.SHAPE_DT), np.array([c]))
    assert tot == m[0]["count"]
    return c, m[0]


def _points(m):
    k = int(m["count"])
    return np.stack([m["px"][:k], m["py"][:k], m["pz"][:k]], axis=1).astype(np.float64), m["depth"][:k].astype(np.float64)


def test_sphere_pairs_keep_the_single_point():
    a, b = O.xf((0, 0, 0)), O.xf((0.9, 0, 0))
    crossed = O.xf((0.2, 0, 0.8), O.axis_angle((0, 0, 1), 0.7))
    for sa, sb, tb in ((O.sphere(0.5), O.sphere(0.5), b), (O.sphere(0.5), O.box(0.5, 0.5, 0.5), b),
                       (O.box(0.5, 0.5, 0.5), O.sphere(0.5), b), (O.capsule(0.45, 1.0), O.sphere(0.5), b),
                       (O.capsule(0.45, 1.0), O.capsule(0.45, 0.5), crossed)):
        c, m = _manifold(a, sa, tb, sb)
        assert m["count"] == 1
        assert (m["px"][0], m["py"][0], m["pz"][0], m["depth"][0]) == (c["px"], c["py"], c["pz"], c["depth"])
        assert (m["nx"], m["ny"], m["nz"]) == (c["nx"], c["ny"], c["nz"])


def test_stacked_boxes_give_the_overlap_rectangle():
    # unit cube on a 2x2x1 slab, shifted so the overlap rectangle is [0.25,1]x[-0.5,0.5], 0.1 deep
    a = O.xf((0, 0, 0))
    b = O.xf((0.75, 0, 0.9))
    c, m = _manifold(a, O.box(1.0, 1.0, 0.5), b, O.box(0.5, 0.5, 0.5))
    np.testing.assert_allclose([c["nx"], c["ny"], c["nz"]], [0, 0, 1], atol=1e-5)
    p, d = _points(m)
    assert m["count"] == 4
    np.testing.assert_allclose(d, 0.1, atol=1e-5)
    np.testing.assert_allclose(p[:, 2], 0.45, atol=1e-5)   # midway between z=0.4 (B's bottom) and z=0.5 (A's top)
    got = sorted((round(x, 4), round(y, 4)) for x, y in p[:, :2])
    assert got == [(0.25, -0.5), (0.25, 0.5), (1.0, -0.5), (1.0, 0.5)]


def test_yawed_stack_is_reduced_to_four_points_of_the_octagon():
    a = O.xf((0, 0, 0))
    b = O.xf((0, 0, 0.95), O.axis_angle((0, 0, 1), np.pi / 4))
    c, m = _manifold(a, O.box(0.5, 0.5, 0.5), b, O.box(0.5, 0.5, 0.5))
    p, d = _points(m)
    assert m["count"] == 4
    np.testing.assert_allclose(d, 0.05, atol=1e-5)
    # every point is a vertex of the octagon: on the boundary of both squares' intersection
    r = np.sqrt(0.5) * 0.5
    for x, y in p[:, :2]:
        on_a = np.isclose(max(abs(x), abs(y)), 0.5, atol=1e-5)
        on_b = np.isclose((abs(x) + abs(y)) * np.sqrt(0.5), 0.5, atol=1e-5)
        assert on_a and on_b, (x, y, r)
    # spread: the four points span the patch (area of their hull well above a sliver)
    q = p[:, :2] - p[:, :2].mean(axis=0)
    assert np.linalg.matrix_rank(q, tol=1e-3) == 2


# ---- independent float64 restatement: clip in the reference face's 2-D coordinates ------------------
def _rot(q):
    x, y, z, w = [float(v) for v in q]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _clip2d(poly, lo, hi):
    """Sutherland-Hodgman of a 2-D polygon against the rectangle [lo, hi] (float64)."""
    for axis in (0, 1):
        for sign, bound in ((1.0, hi[axis]), (-1.0, -lo[axis])):
            out = []
            for k in range(len(poly)):
                p, q = poly[k - 1], poly[k]
                dp, dq = sign * p[axis] - bound, sign * q[axis] - bound
                if (dp <= 0) != (dq <= 0):
                    out.append(p + (q - p) * (dp / (dp - dq)))
                if dq <= 0:
                    out.append(q)
            poly = out
            if not poly:
                return poly
    return poly


def _numpy_candidates(c, xa, sa, xb, sb):
    """All clipped vertices at or below the reference face (before reduction): (positions, depths)."""
    n = np.array([c["nx"], c["ny"], c["nz"]], dtype=np.float64)
    fr = []
    for x, s in ((xa, sa), (xb, sb)):
        x = x.astype(np.float64)
        fr.append((x[:3], _rot(x[3:7]), np.abs(np.array(s[1:4], dtype=np.float64) * x[7:10])))
    al = [np.abs(n @ f[1]) for f in fr]
    ref_is_a = al[0].max() >= al[1].max()
    (cr, Rr, hr), (ci, Ri, hi_) = (fr[0], fr[1]) if ref_is_a else (fr[1], fr[0])
    ndir = n if ref_is_a else -n
    i = int(np.argmax(np.abs(ndir @ Rr)))
    nr = Rr[:, i] * np.sign(ndir @ Rr[:, i])
    j = int(np.argmax(np.abs(nr @ Ri)))
    fcen = ci - Ri[:, j] * np.sign(nr @ Ri[:, j]) * hi_[j]
    u, v = (j + 1) % 3, (j + 2) % 3
    quad = [fcen + su * Ri[:, u] * hi_[u] + sv * Ri[:, v] * hi_[v] for su, sv in ((1, 1), (-1, 1), (-1, -1), (1, -1))]
    wu, wv = (i + 1) % 3, (i + 2) % 3
    # coordinates (s, t, height) in the reference box's frame
    loc = [np.array([(p - cr) @ Rr[:, wu], (p - cr) @ Rr[:, wv], (p - cr) @ nr]) for p in quad]
    out = _clip2d(loc, np.array([-hr[wu], -hr[wv]]), np.array([hr[wu], hr[wv]]))
    pos, dep = [], []
    for q in out:
        sep = q[2] - hr[i]
        if sep <= 0:
            w = cr + Rr[:, wu] * q[0] + Rr[:, wv] * q[1] + nr * (q[2] - 0.5 * sep)
            pos.append(w)
            dep.append(-sep)
    return np.array(pos).reshape(-1, 3), np.array(dep)


def _box_sdist(p, x, s):
    """Signed-ish distance of world point p to the box (<= 0 inside), float64."""
    x = x.astype(np.float64)
    loc = (p - x[:3]) @ _rot(x[3:7])
    h = np.abs(np.array(s[1:4], dtype=np.float64) * x[7:10])
    q = np.abs(loc) - h
    return np.linalg.norm(np.maximum(q, 0.0)) + min(q.max(), 0.0)


def test_random_box_pairs_against_the_numpy_clipper():
    rng = np.random.default_rng(7)
    multi = reduced = 0
    for _ in range(600):
        qa = rng.normal(size=4); qa /= np.linalg.norm(qa)
        qb = rng.normal(size=4); qb /= np.linalg.norm(qb)
        if rng.random() < 0.4:   # near-parallel faces: the interesting manifolds
            qb = qa.copy()
            if rng.random() < 0.5:
                qb = qb + rng.normal(size=4) * 0.02
                qb /= np.linalg.norm(qb)
        sa = O.box(*rng.uniform(0.25, 0.6, 3))
        sb = O.box(*rng.uniform(0.25, 0.6, 3))
        xa = O.xf(rng.uniform(-1, 1, 3), qa, rng.uniform(0.8, 1.3, 3))
        xb = O.xf(xa[:3] + rng.uniform(-0.6, 0.6, 3), qb)
        hit, c, _, _ = O.collide_pair(xa, sa, xb, sb)
        if not hit:
            continue
        c = c.copy(); c["a"], c["b"] = 0, 1
        m, _ = O.manifolds(np.stack([xa, xb]), np.array([sa, sb], dtype=O.SHAPE_DT), np.array([c]))
        p, d = _points(m[0])
        cand_p, cand_d = _numpy_candidates(c, xa, sa, xb, sb)
        if len(cand_p) == 0:
            assert m[0]["count"] == 1
            continue
        # every reported point is one of the independently clipped vertices, with its depth
        assert 1 <= len(p) <= min(4, len(cand_p))
        for pk, dk in zip(p, d):
            e = np.linalg.norm(cand_p - pk, axis=1)
            k = int(np.argmin(e))
            assert e[k] < 2e-4, (pk, cand_p)
            assert abs(cand_d[k] - dk) < 2e-4
        if len(cand_p) <= 4:
            assert len(p) == len(cand_p)
        else:
            reduced += 1
            assert len(p) == 4
            assert abs(d.max() - cand_d.max()) < 2e-4   # the deepest vertex survives the reduction
        multi += len(p) > 1
        # geometry: each point lies within half its depth of both boxes
        for pk, dk in zip(p, d):
            assert _box_sdist(pk, xa, sa) <= 0.5 * dk + 2e-4
            assert _box_sdist(pk, xb, sb) <= 0.5 * dk + 2e-4
    assert multi > 100 and reduced > 10, (multi, reduced)


def test_scene_totals_and_determinism():
    import axcd
    s = axcd.config_scene("C0")
    rc, bb = O.refit(s.xf, s.shapes, s.hull)
    pairs = O.broadphase(bb, brute=True)
    con, _, _ = O.narrowphase(s.xf, s.shapes, pairs, s.hull)
    m1, t1 = O.manifolds(s.xf, s.shapes, con)
    m2, t2 = O.manifolds(s.xf, s.shapes, con, nthreads=4)
    assert m1.tobytes() == m2.tobytes() and t1 == t2 == int(m1["count"].sum())
    assert np.array_equal(m1["a"], con["a"]) and np.array_equal(m1["b"], con["b"])
    bb_mask = (s.shapes["type"][con["a"]] == 1) & (s.shapes["type"][con["b"]] == 1)
    assert (m1["count"][~bb_mask] == 1).all()
    assert (m1["count"] >= 1).all() and (m1["count"] <= 4).all()
    assert (m1["count"][bb_mask] > 1).any()


# ---- capsule against box --------------------------------------------------------------------------------
def _seg_dist(p, x, sh):
    """Distance (float64) from world point p to the capsule's core segment."""
    x = x.astype(np.float64)
    e = _rot(x[3:7])[:, 1] * (float(sh[2]) * 0.5 * x[8])
    a, b = x[:3] - e, x[:3] + e
    u = np.clip((p - a) @ (b - a) / max((b - a) @ (b - a), 1e-300), 0, 1)
    return np.linalg.norm(p - a - u * (b - a))


def test_capsule_lying_on_a_box_gives_two_points():
    box, cap = O.box(2.0, 2.0, 0.5), O.capsule(0.3, 2.0)
    a = O.xf((0, 0, 0))
    b = O.xf((0.2, 0.1, 0.75))                      # axis along y, 0.05 deep into the top face
    c, m = _manifold(a, box, b, cap)
    np.testing.assert_allclose([c["nx"], c["ny"], c["nz"]], [0, 0, 1], atol=1e-4)
    p, d = _points(m)
    assert m["count"] == 2
    np.testing.assert_allclose(d, 0.05, atol=1e-5)
    np.testing.assert_allclose(p, [[0.2, -0.9, 0.475], [0.2, 1.1, 0.475]], atol=1e-5)
    # the same pair with the capsule as body a: same points, normal flipped
    c2, m2 = _manifold(b, cap, a, box)
    p2, d2 = _points(m2)
    assert m2["count"] == 2
    np.testing.assert_allclose(sorted(map(tuple, np.round(p2, 5))), sorted(map(tuple, np.round(p, 5))), atol=1e-5)
    np.testing.assert_allclose([c2["nz"]], [-1], atol=1e-4)
    # overhanging the edge: the segment is clipped at the face boundary y = 2
    b3 = O.xf((0.2, 1.8, 0.75))
    _, m3 = _manifold(a, box, b3, cap)
    p3, _ = _points(m3)
    assert m3["count"] == 2
    np.testing.assert_allclose(sorted(p3[:, 1]), [0.8, 2.0], atol=1e-5)
    # tilted: only the lower end is within reach of the face
    b4 = O.xf((0.0, 0.0, 1.2), O.axis_angle((1, 0, 0), np.pi / 6))
    c4, m4 = _manifold(a, box, b4, cap)
    assert m4["count"] == 1
    p4, d4 = _points(m4)
    assert abs(d4[0] - c4["depth"]) < 1e-3


def test_random_capsule_box_manifolds_lie_on_both_shapes():
    rng = np.random.default_rng(11)
    two = one = 0
    for _ in range(500):
        qa = rng.normal(size=4); qa /= np.linalg.norm(qa)
        box = O.box(*rng.uniform(0.4, 1.0, 3))
        cap = O.capsule(rng.uniform(0.15, 0.3), rng.uniform(0.4, 1.5))
        xa = O.xf(rng.uniform(-1, 1, 3), qa, rng.uniform(0.8, 1.3, 3))
        R = _rot(qa)
        # put the capsule near a face, its axis roughly in the face plane
        k = rng.integers(0, 3)
        h = abs(box[1 + k] * xa[7 + k])
        axis = R[:, (k + 1) % 3] * np.cos(0.2 * rng.normal()) + R[:, k] * 0.15 * rng.normal() + R[:, (k + 2) % 3] * rng.normal() * 0.5
        axis /= np.linalg.norm(axis)
        y = np.array([0.0, 1.0, 0.0])
        v = np.cross(y, axis); sn = np.linalg.norm(v); cs = y @ axis
        qb = np.array([*(v / max(sn, 1e-12) * np.sin(np.arctan2(sn, cs) / 2)), np.cos(np.arctan2(sn, cs) / 2)])
        pos = xa[:3] + R[:, k] * (h + cap[1] * rng.uniform(0.6, 1.0)) * rng.choice([-1, 1]) + R[:, (k + 1) % 3] * rng.uniform(-0.3, 0.3)
        xb = O.xf(pos, qb)
        for (x0, s0, x1, s1) in ((xa, box, xb, cap), (xb, cap, xa, box)):
            hit, c, _, _ = O.collide_pair(x0, s0, x1, s1)
            if not hit:
                continue
            c = c.copy(); c["a"], c["b"] = 0, 1
            m, _ = O.manifolds(np.stack([x0, x1]), np.array([s0, s1], dtype=O.SHAPE_DT), np.array([c]))
            p, d = _points(m[0])
            assert 1 <= len(p) <= 2
            two += len(p) == 2
            one += len(p) == 1
            if len(p) == 1 and np.allclose(p[0], [c["px"], c["py"], c["pz"]]):
                continue          # fallback to the narrowphase point
            xc, sc = (x1, s1) if s1[0] == 2 else (x0, s0)
            xbx, sbx = (x0, s0) if s1[0] == 2 else (x1, s1)
            for pk, dk in zip(p, d):
                # half a depth inside the capsule's surface, and no farther than that from the box
                # (the end of the clipped segment above the point is at r - dk/2; a tilted segment may pass
                # closer, never farther)
                assert _seg_dist(pk, xc, sc) <= sc[1] - 0.5 * dk + 3e-4, (pk, dk)
                assert _box_sdist(pk, xbx, sbx) <= 3e-4
                assert dk <= 1.5 * c["depth"] + 2e-3     # measured along the face normal, not the (minimal) contact normal
    assert two > 60 and one > 60, (two, one)


# ------------------------------------------------------------------ capsule against capsule -----------
def test_parallel_capsules_give_the_two_ends_of_the_overlap():
    capA, capB = O.capsule(0.3, 2.0), O.capsule(0.25, 1.0)
    a = O.xf((0, 0, 0))
    b = O.xf((0.5, 0.8, 0))                           # both along y; B's segment spans y in [0.3, 1.3], A's [-1, 1]
    c, m = _manifold(a, capA, b, capB)
    np.testing.assert_allclose([c["nx"], c["ny"], c["nz"]], [1, 0, 0], atol=1e-5)
    p, d = _points(m)
    assert m["count"] == 2
    np.testing.assert_allclose(d, 0.05, atol=1e-6)    # 0.3 + 0.25 - 0.5
    np.testing.assert_allclose(p, [[0.275, 0.3, 0], [0.275, 1.0, 0]], atol=1e-6)
    # the same pair the other way round: same points, normal flipped
    c2, m2 = _manifold(b, capB, a, capA)
    p2, d2 = _points(m2)
    assert m2["count"] == 2 and c2["nx"] < -0.999
    np.testing.assert_allclose(sorted(map(tuple, np.round(p2, 5))), sorted(map(tuple, np.round(p, 5))), atol=1e-5)
    np.testing.assert_allclose(d2, 0.05, atol=1e-6)
    # end to end along the common axis: no stretch in common -> the narrowphase point
    b3 = O.xf((0.0, 1.9, 0))
    c3, m3 = _manifold(a, capA, b3, capB)
    assert m3["count"] == 1 and (m3["px"][0], m3["py"][0], m3["pz"][0]) == (c3["px"], c3["py"], c3["pz"])
    # a small tilt (2 degrees, inside the 5.7 degree window) about z: B's far end rises out of reach -> one point, at the near end
    b4 = O.xf((0.53, 0.8, 0), O.axis_angle((0, 0, 1), -np.radians(2.0)))
    c4, m4 = _manifold(a, capA, b4, capB)
    p4, d4 = _points(m4)
    assert 1 <= m4["count"] <= 2 and d4.max() <= c4["depth"] + 2e-3 and d4.min() < 0.6 * d4.max()
    # crossed at 40 degrees: single point
    b5 = O.xf((0.5, 0.0, 0), O.axis_angle((1, 0, 0), 0.7))
    _, m5 = _manifold(a, capA, b5, capB)
    assert m5["count"] == 1


def test_random_near_parallel_capsule_manifolds_lie_in_both_capsules():
    rng = np.random.default_rng(17)
    two = one = 0
    for _ in range(600):
        qa = rng.normal(size=4); qa /= np.linalg.norm(qa)
        sa, sb = O.capsule(rng.uniform(0.15, 0.3), rng.uniform(0.6, 2.0)), O.capsule(rng.uniform(0.15, 0.3), rng.uniform(0.6, 2.0))
        xa = O.xf(rng.uniform(-1, 1, 3), qa, rng.uniform(0.8, 1.3, 3))
        R = _rot(qa)
        # B: A's rotation composed with a tilt of 0..8 degrees, beside A at just under the sum of the radii
        tilt = O.axis_angle(rng.normal(size=3), np.radians(rng.uniform(0, 8)))
        qb = O.quat_mul(tilt, qa.astype(np.float32)) if hasattr(O, "quat_mul") else None
        if qb is None:
            pytest.skip("oracle_lib has no quat_mul")
        side = R[:, 0] * np.cos(t := rng.uniform(0, 2 * np.pi)) + R[:, 2] * np.sin(t)
        pos = xa[:3] + side * (sa[1] + sb[1]) * rng.uniform(0.7, 0.98) + R[:, 1] * rng.uniform(-0.8, 0.8)
        xb = O.xf(pos, qb, rng.uniform(0.8, 1.3, 3))
        hit, c, _, _ = O.collide_pair(xa, sa, xb, sb)
        if not hit:
            continue
        c = c.copy(); c["a"], c["b"] = 0, 1
        m, _ = O.manifolds(np.stack([xa, xb]), np.array([sa, sb], dtype=O.SHAPE_DT), np.array([c]))
        p, d = _points(m[0])
        assert 1 <= len(p) <= 2
        two += len(p) == 2
        one += len(p) == 1
        if len(p) == 1 and np.allclose(p[0], [c["px"], c["py"], c["pz"]]):
            continue
        n = np.array([c["nx"], c["ny"], c["nz"]], dtype=np.float64)
        for pk, dk in zip(p, d):
            # the point is the midpoint of the overlap along n: inside both capsules, about r - depth/2 from each
            # axis (exactly so for parallel axes; a tilt shifts it sideways by a fraction of the depth), and no
            # depth exceeds what the narrowphase found by much
            for xx, ss in ((xa, sa), (xb, sb)):
                dist = _seg_dist(pk, xx, ss)
                assert dist <= ss[1] + 0.01, (pk, dk)
                assert abs(dist - (ss[1] - 0.5 * dk)) <= 0.02, (pk, dk, dist)
            assert dk <= 1.1 * c["depth"] + 2e-3     # measured along the shared normal, not along pB - pA
        assert abs(d.max() - c["depth"]) < 0.05          # one end is (close to) the deepest point
    assert two > 80 and one > 40, (two, one)
