"""Validates the oracle's scene queries against independent float64 numpy:
AABB overlap query vs a vectorised closed-interval test; ray casts vs a root search on the signed
distance of each shape (no closed forms shared with the oracle)."""
import numpy as np
import pytest

import axcd
import oracle_lib as O


def _scene(n=400, seed=11, L=6.0, capsules=True, hulls=False):
    rng = np.random.default_rng(seed)
    s = axcd.generate_scene(n, seed, L, frac_box=0.4, frac_sphere=0.3 if hulls else 0.6)
    if capsules:
        k = np.where(s.shapes["type"] == 0)[0][::2]
        s.shapes["type"][k] = 2
        s.shapes["p0"][k] = rng.uniform(0.15, 0.3, len(k))
        s.shapes["p1"][k] = rng.uniform(0.3, 0.9, len(k))
    s.xf[:, 7:10] = rng.uniform(0.7, 1.4, (n, 3)).astype(np.float32)
    return s


def test_aabb_query_matches_vectorised_numpy():
    s = _scene()
    _, bb = O.refit(s.xf, s.shapes, s.hull)
    rng = np.random.default_rng(3)
    c = rng.uniform(0, 6, (64, 3))
    h = rng.uniform(0.05, 1.5, (64, 3))
    q = np.concatenate([c - h, c + h], axis=1).astype(np.float32)
    q[0] = bb[5]                       # touching / identical boxes count (closed intervals)
    q[1, :3], q[1, 3:] = bb[7, 3:], bb[7, 3:] + 1     # min corner == a body's max corner
    got = O.query_aabbs(bb, q)
    ov = ((q[:, None, :3] <= bb[None, :, 3:]) & (q[:, None, 3:] >= bb[None, :, :3])).all(axis=2)
    exp = np.argwhere(ov).astype(np.uint32)
    assert np.array_equal(got, exp)
    assert (got[got[:, 0] == 1][:, 1] == 7).any()


def _rot(q):
    x, y, z, w = [float(v) for v in q]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def _sdf(p, x, sh):
    """Signed distance (float64) from points p (k,3) to the shape."""
    x = x.astype(np.float64)
    m = p - x[:3]
    t = int(sh["type"])
    if t == 0:
        return np.linalg.norm(m, axis=1) - float(sh["p0"])
    R = _rot(x[3:7])
    if t == 1:
        h = np.abs(np.array([sh["p0"], sh["p1"], sh["p2"]], dtype=np.float64) * x[7:10])
        q = np.abs(m @ R) - h
        return np.linalg.norm(np.maximum(q, 0), axis=1) + np.minimum(q.max(axis=1), 0)
    e = R[:, 1] * (float(sh["p1"]) * 0.5 * x[8])
    a, b = -e, e
    ba = b - a
    u = np.clip(((m - a) @ ba) / max(ba @ ba, 1e-300), 0, 1)
    return np.linalg.norm(m - a - u[:, None] * ba, axis=1) - float(sh["p0"])


def _first_root(o, d, tmax, x, sh):
    """First t in [0,tmax] with sdf <= 0: dense sampling + bisection (float64)."""
    ts = np.linspace(0.0, tmax, 4001)
    f = _sdf(o[None] + ts[:, None] * d[None], x, sh)
    idx = np.where(f <= 0)[0]
    if len(idx) == 0:
        return None
    k = idx[0]
    if k == 0:
        return 0.0
    lo, hi = ts[k - 1], ts[k]
    for _ in range(60):
        mid = 0.5 * (lo + hi)
        if _sdf((o + mid * d)[None], x, sh)[0] <= 0:
            hi = mid
        else:
            lo = mid
    return hi


def test_ray_single_shapes_against_sdf_root_search():
    s = _scene(n=120, seed=5, L=4.0)
    _, bb = O.refit(s.xf, s.shapes, s.hull)
    rng = np.random.default_rng(9)
    checked = {0: 0, 1: 0, 2: 0}
    for i in range(s.n):
        c = s.xf[i, :3].astype(np.float64)
        o = c + rng.normal(size=3) * 2.0
        tgt = c + rng.normal(size=3) * 0.25
        d = tgt - o
        d /= np.linalg.norm(d)
        rays = O.make_rays([o], [d], 10.0)
        hit = O.raycast(s.xf[i:i + 1], s.shapes[i:i + 1], bb[i:i + 1], rays)[0]
        ref = _first_root(o.astype(np.float32).astype(np.float64), d.astype(np.float32).astype(np.float64), 10.0, s.xf[i], s.shapes[i])
        if ref is None:
            # grazing rays may differ by rounding; a reported hit must then be a near miss
            if hit["body"] != O.NO_HIT:
                p = o + hit["t"] * d
                assert abs(_sdf(p[None], s.xf[i], s.shapes[i])[0]) < 1e-3
            continue
        assert hit["body"] == 0, (i, s.shapes[i], ref)
        assert abs(hit["t"] - ref) < 2e-4 * max(1.0, ref), (i, s.shapes[i], hit, ref)
        if ref > 0:
            # the normal is the sdf gradient at the hit point
            p = o + float(hit["t"]) * d
            eps = 1e-5
            g = np.array([(_sdf((p + eps * e)[None], s.xf[i], s.shapes[i])[0] -
                           _sdf((p - eps * e)[None], s.xf[i], s.shapes[i])[0]) / (2 * eps) for e in np.eye(3)])
            n = np.array([hit["nx"], hit["ny"], hit["nz"]], dtype=np.float64)
            if np.linalg.norm(g) > 0.99:   # away from edges, where the gradient is not unique
                assert np.dot(n, g) > 0.98, (i, s.shapes[i], n, g)
        checked[int(s.shapes["type"][i])] += 1
    assert all(v > 10 for v in checked.values()), checked


def test_ray_closest_body_in_a_scene():
    s = _scene(n=300, seed=21, L=5.0)
    _, bb = O.refit(s.xf, s.shapes, s.hull)
    rng = np.random.default_rng(2)
    o = rng.uniform(-1, 6, (60, 3))
    d = rng.normal(size=(60, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[0] = (1, 0, 0)   # axis-parallel rays exercise the zero-component rule
    d[1] = (0, -1, 0)
    rays = O.make_rays(o, d, 8.0)
    hits = O.raycast(s.xf, s.shapes, bb, rays)
    nhit = 0
    for q in range(len(rays)):
        oq = np.array([rays[q][k] for k in ("ox", "oy", "oz")], dtype=np.float64)
        dq = np.array([rays[q][k] for k in ("dx", "dy", "dz")], dtype=np.float64)
        best, bi = None, None
        for i in range(s.n):
            r = _first_root(oq, dq, 8.0, s.xf[i], s.shapes[i]) if np.linalg.norm(np.cross(s.xf[i, :3] - oq, dq)) < 2.5 else None
            if r is not None and (best is None or r < best):
                best, bi = r, i
        if best is None:
            assert hits[q]["body"] == O.NO_HIT or abs(hits[q]["t"]) <= 8.0
            continue
        nhit += 1
        assert hits[q]["body"] == bi or abs(hits[q]["t"] - best) < 1e-3, (q, hits[q], best, bi)
        assert abs(hits[q]["t"] - best) < 1e-3
    assert nhit > 20


def test_ray_miss_world_filter_and_exact_hull_hits():
    s = axcd.generate_scene(50, 4, 4.0, frac_box=0.3, frac_sphere=0.3)     # 40 % hulls
    _, bb = O.refit(s.xf, s.shapes, s.hull)
    rays = O.make_rays([[-50, -50, -50], [2, 2, -5]], [[-1, 0, 0], [0, 0, 1]], 100.0)
    hits = O.raycast(s.xf, s.shapes, bb, rays, hull=s.hull)
    assert hits[0]["body"] == O.NO_HIT and hits[0]["t"] == np.float32(100.0)
    # world filter: bodies of the other world are invisible
    wid = (np.arange(s.n) % 2).astype(np.uint32)
    r0 = O.make_rays([[2, 2, -5]] * 2, [[0, 0, 1]] * 2, 100.0)
    r0["world"] = [0, 1]
    h = O.raycast(s.xf, s.shapes, bb, r0, world_id=wid, hull=s.hull)
    for k in (0, 1):
        if h[k]["body"] != O.NO_HIT:
            assert wid[h[k]["body"]] == k
    # hull bodies: exact hits by conservative advancement, checked against the hull's half-spaces
    from scipy.spatial import ConvexHull
    hid = np.where(s.shapes["type"] == 4)[0]
    checked = 0
    for i in hid[:12]:
        first, cnt = np.array([s.shapes["p0"][i]], np.float32).view(np.uint32)[0], np.array([s.shapes["p1"][i]], np.float32).view(np.uint32)[0]
        x = s.xf[i].astype(np.float64)
        R = _rot(x[3:7])
        pts = (s.hull[first:first + cnt].astype(np.float64) * x[7:10]) @ R.T + x[:3]
        eq = ConvexHull(pts).equations            # n.x + d <= 0 inside
        o = x[:3] + np.array([0.06, -0.04, -6.0])
        d = np.array([0.0, 0.0, 1.0])
        num, den = -(eq[:, :3] @ o + eq[:, 3]), eq[:, :3] @ d
        t_in = max([num[k] / den[k] for k in range(len(eq)) if den[k] < 0] + [0.0])
        t_out = min([num[k] / den[k] for k in range(len(eq)) if den[k] > 0] + [1e30])
        h = O.raycast(s.xf[i:i + 1], s.shapes[i:i + 1], bb[i:i + 1], O.make_rays([o], [d], 50.0), hull=s.hull)[0]
        if t_in <= t_out:
            assert h["body"] == 0 and h["flags"] == 0
            assert -3e-4 < t_in - h["t"] < 3e-4, (t_in, h)       # stops when the gap is <= 1e-4
            checked += 1
        else:
            assert h["body"] == O.NO_HIT
    assert checked >= 5
