"""Cylinder shapes (gui::ShapeType::Cylinder, include/axiom/gui/body_inspector.hpp:24).

Oracle validation on the CPU against closed forms and an independent QP over a dense sampling of the cylinder's
surface; the CUDA path against the oracle, bit for bit (refit, pair set, contacts), in the GPU test."""
import numpy as np
import pytest
from scipy.optimize import minimize

import axcd
import oracle_lib as O

TOL = 2e-4


def rot_matrix(q):
    x, y, z, w = [float(v) for v in q]
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - w * z), 2 * (x * z + w * y)],
                     [2 * (x * y + w * z), 1 - 2 * (x * x + z * z), 2 * (y * z - w * x)],
                     [2 * (x * z - w * y), 2 * (y * z + w * x), 1 - 2 * (x * x + y * y)]])


def cyl_points(t, r, h, n_ring=48):
    """Both rims of the cylinder sampled at n_ring points each (circumscribed polygon radius, so the sampled hull
    CONTAINS the cylinder and differs from it by r (1/cos(pi/n) - 1)), world space, float64."""
    ang = np.linspace(0, 2 * np.pi, n_ring, endpoint=False)
    rr = float(r) / np.cos(np.pi / n_ring)
    ring = np.stack([rr * np.cos(ang), np.zeros_like(ang), rr * np.sin(ang)], axis=1)
    loc = np.concatenate([ring + [0, float(h) / 2, 0], ring - [0, float(h) / 2, 0]])
    t = np.asarray(t, float)
    return (rot_matrix(t[3:7]) @ (loc * t[7:10]).T).T + t[:3]


def qp_distance(PA, PB):
    """Distance between conv(PA) and conv(PB) (vertex-based QP, as tests/test_oracle_narrow.py)."""
    na, nb = len(PA), len(PB)

    def f(x):
        d = x[:na] @ PA - x[na:] @ PB
        return d @ d

    def g(x):
        d = x[:na] @ PA - x[na:] @ PB
        return np.concatenate([2 * PA @ d, -2 * PB @ d])
    cons = [{"type": "eq", "fun": lambda x: x[:na].sum() - 1, "jac": lambda x: np.concatenate([np.ones(na), np.zeros(nb)])},
            {"type": "eq", "fun": lambda x: x[na:].sum() - 1, "jac": lambda x: np.concatenate([np.zeros(na), np.ones(nb)])}]
    x0 = np.concatenate([np.full(na, 1 / na), np.full(nb, 1 / nb)])
    res = minimize(f, x0, jac=g, bounds=[(0, 1)] * (na + nb), constraints=cons, method="SLSQP",
                   options={"maxiter": 400, "ftol": 1e-14})
    return float(np.sqrt(max(res.fun, 0.0)))


def test_cylinder_refit_is_the_exact_box():
    rng = np.random.default_rng(1)
    for _ in range(200):
        r, h = rng.uniform(0.1, 1.0), rng.uniform(0.1, 2.0)
        t = O.xf(rng.uniform(-5, 5, 3), O.axis_angle(rng.normal(size=3), rng.uniform(0, 6.28)), rng.uniform(0.5, 2.0, 3))
        rc, bb = O.refit(t.reshape(1, 10), np.array([O.cylinder(r, h)], O.SHAPE_DT))
        assert rc == 0
        pts = cyl_points(t, np.float32(r), np.float32(h), 4000)      # polygon error 3e-7 r
        np.testing.assert_allclose(bb[0, :3], pts.min(axis=0), atol=2e-5)
        np.testing.assert_allclose(bb[0, 3:], pts.max(axis=0), atol=2e-5)


def test_cylinder_sphere_closed_forms():
    """Sphere beside the wall: gap = x - r_c - r_s; sphere above the cap (inside the rim): gap = y - h/2 - r_s, contact
    position under the sphere, not at the cap centre."""
    c = O.cylinder(0.5, 2.0)
    for x in (1.0, 0.76, 0.74, 0.6):
        hit, con, dist, epa = O.collide_pair(O.xf(), c, O.xf((x, 0.3, 0)), O.sphere(0.25))
        assert abs(dist - (x - 0.75)) < TOL and hit == (x < 0.75)
        if hit:
            np.testing.assert_allclose([con["nx"], con["ny"], con["nz"]], (1, 0, 0), atol=1e-3)
    for y in (1.5, 1.26, 1.2):
        hit, con, dist, epa = O.collide_pair(O.xf(), c, O.xf((0.2, y, 0.1)), O.sphere(0.25))
        assert abs(dist - (y - 1.25)) < TOL and hit == (y < 1.25)
        if hit:
            np.testing.assert_allclose([con["nx"], con["ny"], con["nz"]], (0, 1, 0), atol=1e-3)
            np.testing.assert_allclose([con["px"], con["pz"]], (0.2, 0.1), atol=1e-3)


def test_crossed_and_stacked_cylinders():
    c = O.cylinder(0.5, 2.0)
    qx = O.axis_angle((1, 0, 0), np.pi / 2)
    hit, con, dist, epa = O.collide_pair(O.xf(), c, O.xf((0.8, 0, 0), qx), c)     # axes Y and Z, 0.8 apart along x
    assert hit and epa and abs(con["depth"] - 0.2) < 1e-3
    np.testing.assert_allclose([con["nx"], con["ny"], con["nz"]], (1, 0, 0), atol=2e-3)
    hit, con, dist, epa = O.collide_pair(O.xf(), c, O.xf((0.1, 1.9, 0.05)), c)     # stacked cap on cap, 0.1 deep
    assert hit and abs(con["depth"] - 0.1) < 1e-3
    np.testing.assert_allclose([con["nx"], con["ny"], con["nz"]], (0, 1, 0), atol=2e-3)
    hit, con, dist, epa = O.collide_pair(O.xf(), c, O.xf((1.2, 0, 0), qx), c)
    assert not hit and abs(dist - 0.2) < TOL


def test_random_cylinder_pairs_vs_qp():
    """Cylinder against box / cylinder / sphere in random poses with non-uniform scales: GJK distance against the
    vertex QP over a dense surface sampling; for penetrating pairs, shifting B by depth along the normal must leave
    the shapes (nearly) touching and a slightly smaller shift must not."""
    rng = np.random.default_rng(7)
    checked = 0
    for it in range(40):
        ra, ha = rng.uniform(0.25, 0.5), rng.uniform(0.3, 1.0)
        ta = O.xf(rng.uniform(0, 0.4, 3), O.axis_angle(rng.normal(size=3), rng.uniform(0, 6.28)), rng.uniform(0.7, 1.4, 3))
        tb_pos = rng.uniform(0, 0.4, 3) + rng.normal(size=3) * 0.5
        tb = O.xf(tb_pos, O.axis_angle(rng.normal(size=3), rng.uniform(0, 6.28)), rng.uniform(0.7, 1.4, 3))
        PA = cyl_points(ta, np.float32(ra), np.float32(ha))
        kind = it % 2
        if kind == 0:
            rb, hb = rng.uniform(0.25, 0.5), rng.uniform(0.3, 1.0)
            sb = O.cylinder(rb, hb)
            PB = cyl_points(tb, np.float32(rb), np.float32(hb))
        else:
            hbx = rng.uniform(0.2, 0.5, 3)
            sb = O.box(*hbx)
            corners = np.array([[sx, sy, sz] for sx in (-1, 1) for sy in (-1, 1) for sz in (-1, 1)], np.float32) * np.float32(hbx)
            PB = (rot_matrix(tb[3:7]) @ (corners.astype(float) * tb[7:10].astype(float)).T).T + tb[:3].astype(float)
        hit, con, dist, epa = O.collide_pair(ta, O.cylinder(ra, ha), tb, sb)
        d = qp_distance(PA, PB)
        SLACK = 3.5e-3      # the circumscribed 48-gons stick out of the cylinders by up to 1.5e-3 each
        if d > 1e-2:
            assert not hit and d - 1e-4 <= dist <= d + SLACK, (dist, d)
            checked += 1
        elif hit and con["depth"] > 1e-2:
            n = np.array([con["nx"], con["ny"], con["nz"]], float)
            assert abs(np.linalg.norm(n) - 1) < 1e-4
            d2 = qp_distance(PA, PB + n * (con["depth"] + 8e-3))     # past the contact: apart, by about the extra shift
            assert 8e-3 - SLACK - 5e-4 < d2 < 8e-3 + 5e-4, (d2, con["depth"], con["status"])
            d3 = qp_distance(PA, PB + n * (con["depth"] - 5e-3))     # short of it: still overlapping
            assert d3 < 5e-4
            checked += 1
    assert checked > 20, checked


def cylinder_scene(n=4000, seed=11, domain=16.0):
    s = axcd.generate_scene(n, seed, domain, frac_box=0.35, frac_sphere=0.35)      # 30 % hulls
    rng = np.random.default_rng(seed)
    k = np.where(s.shapes["type"] == axcd.SHAPE_SPHERE)[0][::2]
    s.shapes["type"][k] = axcd.SHAPE_CYLINDER
    s.shapes["p0"][k] = rng.uniform(0.2, 0.45, len(k)).astype(np.float32)
    s.shapes["p1"][k] = rng.uniform(0.3, 0.9, len(k)).astype(np.float32)
    k2 = np.where(s.shapes["type"] == axcd.SHAPE_BOX)[0][::3]
    s.shapes["type"][k2] = axcd.SHAPE_CAPSULE
    s.shapes["p0"][k2] = np.float32(0.2)
    s.shapes["p1"][k2] = np.float32(0.6)
    s.shapes["p2"][k2] = 0.0
    s.xf[:, 7:10] = rng.uniform(0.7, 1.4, (s.n, 3)).astype(np.float32)
    return s


def test_oracle_cylinder_scene_sanity():
    s = cylinder_scene(1500)
    rc, bb = O.refit(s.xf, s.shapes, s.hull)
    assert rc == 0
    pairs = O.broadphase(bb)
    con, _, st = O.narrowphase(s.xf, s.shapes, pairs, s.hull)
    cyl = s.shapes["type"] == axcd.SHAPE_CYLINDER
    involved = cyl[con["a"]] | cyl[con["b"]]
    assert involved.sum() > 100
    n = np.stack([con["nx"], con["ny"], con["nz"]], axis=1)[involved]
    np.testing.assert_allclose(np.linalg.norm(n, axis=1), 1.0, atol=1e-4)
    assert (con["depth"][involved] >= 0).all()
    assert (con["status"][involved] == 302).mean() < 0.2      # curved surfaces may hit EPA's caps, rarely


@pytest.mark.gpu
def test_gpu_cylinder_scene_matches_oracle_bit_for_bit():
    s = cylinder_scene(20000, domain=27.0)
    w = axcd.CollisionWorld.for_scene(s, pairs_per_body=16)
    st = w.step()
    rc, bb = O.refit(s.xf, s.shapes, s.hull, nthreads=8)
    assert np.array_equal(w.aabbs().view(np.uint32), bb.view(np.uint32))
    pairs = O.broadphase(bb, nthreads=8)
    assert np.array_equal(w.pairs(), pairs)
    con, _, ost = O.narrowphase(s.xf, s.shapes, pairs, s.hull, nthreads=8)
    assert st.numContacts == len(con) and st.numPenetrating == ost.numPenetrating
    assert w.contacts().tobytes() == con.tobytes()
    # manifolds (single point for cylinder pairs), a ray cast and a sweep through the same shapes
    w.build_manifolds()
    gm, pts = w.manifolds()
    om, opts = O.manifolds(s.xf, s.shapes, con)
    assert gm.tobytes() == om.tobytes() and pts == opts
    rng = np.random.default_rng(3)
    o = rng.uniform(0, 27, (2000, 3)).astype(np.float32)
    d = rng.normal(size=(2000, 3)).astype(np.float32)
    rays = O.make_rays(o, d, 20.0)
    assert w.raycast(rays).tobytes() == O.raycast(s.xf, s.shapes, bb, rays, hull=s.hull).tobytes()
    cyl = np.where(s.shapes["type"] == axcd.SHAPE_CYLINDER)[0]
    cp = np.stack([cyl[:400], cyl[400:800]], axis=1).astype(np.uint32)
    disp = np.zeros((s.n, 3), np.float32)
    disp[cp[:, 1]] = s.xf[cp[:, 0], :3] - s.xf[cp[:, 1], :3]
    assert w.ccd_pairs(cp, disp).tobytes() == O.ccd_pairs(s.xf, s.shapes, cp, disp, s.hull).tobytes()
    w.close()
