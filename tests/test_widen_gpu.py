"""GPU parity tests for the SURVEY 8(f) "next" rows built on top of the hot path: contact manifolds
(rank 2) and scene queries on the LBVH (rank 4).  Same bar as tests/test_gpu_parity.py: index
results bit-exact against the CPU oracle, floats within REL = 1e-4 relative (and, because both sides
evaluate the same expression trees without contraction, additionally asserted bit-identical)."""
import numpy as np
import pytest

import axcd
import oracle_lib as O

pytestmark = pytest.mark.gpu
REL = 1e-4


def _bits_equal(a, b):
    a, b = np.ascontiguousarray(a), np.ascontiguousarray(b)
    return a.tobytes() == b.tobytes()


def _manifold_parity(s, **kw):
    w = axcd.CollisionWorld.for_scene(s, **kw)
    st = w.step()
    assert w.stats().contactPointCount == 0          # nothing built yet
    w.build_manifolds()
    gm, pts = w.manifolds()
    gc = w.contacts()
    assert len(gm) == st.numContacts == len(gc)
    om, opts = O.manifolds(s.xf, s.shapes, gc, nthreads=8)
    assert np.array_equal(gm["a"], om["a"]) and np.array_equal(gm["b"], om["b"])
    assert np.array_equal(gm["count"], om["count"])                      # integer results: bit-exact
    assert pts == opts == int(gm["count"].sum()) == w.stats().contactPointCount
    for f in ("nx", "ny", "nz"):
        assert _bits_equal(gm[f], om[f])
    for f in ("px", "py", "pz", "depth"):
        np.testing.assert_allclose(gm[f], om[f], rtol=REL, atol=REL * 1e-2)
    bitwise = all(_bits_equal(gm[f], om[f]) for f in ("px", "py", "pz", "depth"))
    w.close()
    return gm, bitwise


def test_manifolds_c0_bit_exact():
    gm, bitwise = _manifold_parity(axcd.config_scene("C0"))
    assert bitwise
    assert (gm["count"] > 1).any() and (gm["count"] <= 4).all()


def _dense_boxes(n=20000, seed=5, L=22.0):
    rng = np.random.default_rng(seed)
    s = axcd.generate_scene(n, seed, L, frac_box=1.0, frac_sphere=0.0)
    # many near-parallel faces (the multi-point manifolds): snap two thirds of the rotations to a
    # common frame, half of those with a small perturbation; non-unit scales
    k = rng.random(n)
    base = np.array([0.0, 0.0, 0.0, 1.0], np.float32)
    q = s.xf[:, 3:7].copy()
    q[k < 0.33] = base
    m = (k >= 0.33) & (k < 0.66)
    qq = base + rng.normal(size=(int(m.sum()), 4)).astype(np.float32) * 0.03
    q[m] = qq / np.linalg.norm(qq, axis=1, keepdims=True)
    s.xf[:, 3:7] = q.astype(np.float32)
    s.xf[:, 7:10] = rng.uniform(0.7, 1.5, (n, 3)).astype(np.float32)
    return s


def test_manifolds_dense_near_parallel_boxes():
    gm, bitwise = _manifold_parity(_dense_boxes(), pairs_per_body=32)
    assert bitwise
    hist = np.bincount(gm["count"], minlength=5)
    assert hist[4] > 500 and hist[2] > 50 and hist[1] > 0, hist


def test_manifolds_mixed_shapes_and_call_order():
    s = axcd.config_scene("C2", scale=0.02)       # boxes / spheres / hulls
    gm, bitwise = _manifold_parity(s)
    assert bitwise
    w = axcd.CollisionWorld.for_scene(s)
    with pytest.raises(axcd.AxcdError) as e:
        w.build_manifolds()                       # no narrowphase yet
    assert e.value.code == 503
    w.step()
    w.build_manifolds()
    with pytest.raises(axcd.AxcdError) as e:
        w.build_manifolds()                       # once per narrowphase
    assert e.value.code == 503
    w.step()
    with pytest.raises(axcd.AxcdError) as e:
        w.manifolds()                             # stale after a new step
    assert e.value.code == 503
    w.close()


def test_manifolds_headline_full_size():
    gm, bitwise = _manifold_parity(axcd.config_scene("headline"))
    assert bitwise
    assert len(gm) == 1_538_646


# ------------------------------------------------------------------ scene queries ---------------
def _query_scene(n=20000, seed=12, L=27.0):
    rng = np.random.default_rng(seed)
    s = axcd.generate_scene(n, seed, L, frac_box=0.4, frac_sphere=0.4)      # 20 % hulls
    k = np.where(s.shapes["type"] == 0)[0][::2]
    s.shapes["type"][k] = 2                                                  # capsules
    s.shapes["p0"][k] = rng.uniform(0.15, 0.3, len(k)).astype(np.float32)
    s.shapes["p1"][k] = rng.uniform(0.3, 0.9, len(k)).astype(np.float32)
    s.xf[:, 7:10] = rng.uniform(0.7, 1.4, (n, 3)).astype(np.float32)
    return s, L


def test_query_aabbs_matches_oracle():
    s, L = _query_scene()
    w = axcd.CollisionWorld.for_scene(s)
    w.update()
    bb = w.aabbs()
    rng = np.random.default_rng(1)
    nq = 3000
    c = rng.uniform(-1, L + 1, (nq, 3))
    h = rng.uniform(0.01, 2.0, (nq, 3)) ** 2
    q = np.concatenate([c - h, c + h], axis=1).astype(np.float32)
    q[0] = bb[17]                                  # identical box
    q[1, :3], q[1, 3:] = bb[23, 3:], bb[23, 3:] + 0.5   # touching at one corner: closed intervals
    q[2] = (-1e9, -1e9, -1e9, 1e9, 1e9, 1e9)       # everything
    q[3] = (5, 5, 5, 4, 6, 6)                      # inverted: nothing
    q[4] = (np.nan, 0, 0, 1, 1, 1)                 # NaN: nothing
    got = w.query_aabbs(q)
    exp = O.query_aabbs(bb, q)
    assert np.array_equal(got, exp)
    assert (got[:, 0] == 2).sum() == s.n and not (got[:, 0] == 3).any() and not (got[:, 0] == 4).any()
    assert ((got[:, 0] == 1) & (got[:, 1] == 23)).any()
    w.close()


def test_raycast_matches_oracle():
    s, L = _query_scene()
    w = axcd.CollisionWorld.for_scene(s)
    w.update()
    bb = w.aabbs()
    rng = np.random.default_rng(4)
    nq = 4000
    o = rng.uniform(-2, L + 2, (nq, 3))
    d = rng.normal(size=(nq, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    d[:300] *= rng.uniform(0.2, 3.0, (300, 1))     # non-unit directions
    d[300] = (1, 0, 0)                              # zero components
    d[301] = (0, 0, -1)
    d[302] = (0, 0, 0)                              # degenerate ray: only "origin inside" hits
    o[303:400] = s.xf[1000:1097, :3]                # origins inside shapes: t = 0
    rays = O.make_rays(o, d, rng.uniform(3.0, 40.0, nq).astype(np.float32))
    got = w.raycast(rays)
    exp = O.raycast(s.xf, s.shapes, bb, rays, hull=s.hull)
    assert np.array_equal(got["body"], exp["body"])                       # bit-exact closest body
    assert np.array_equal(got["flags"], exp["flags"])
    for f in ("t", "nx", "ny", "nz"):
        np.testing.assert_allclose(got[f], exp[f], rtol=REL, atol=REL * 1e-2)
    assert all(_bits_equal(got[f], exp[f]) for f in ("t", "nx", "ny", "nz"))
    hit = got["body"] != axcd.NO_HIT
    assert 0.2 < hit.mean() < 1.0
    assert (got["t"][303:400] == 0).sum() > 80
    kinds = set(np.unique(s.shapes["type"][got["body"][hit]]))
    assert kinds == {0, 1, 2, 4}
    w.close()


def test_queries_in_batched_worlds():
    s = axcd.config_scene("C3", scale=64 / 4096)       # 64 worlds x 256 bodies, all in the same cube
    w = axcd.CollisionWorld.for_scene(s)
    w.update()
    bb = w.aabbs()
    rng = np.random.default_rng(8)
    nq = 1000
    c = rng.uniform(0, 6.35, (nq, 3))
    h = rng.uniform(0.1, 1.5, (nq, 3))
    q = np.concatenate([c - h, c + h], axis=1).astype(np.float32)
    qw = rng.integers(0, 64, nq).astype(np.uint32)
    qw[:5] = 63
    qw[5:10] = 0
    got = w.query_aabbs(q, qw)
    exp = O.query_aabbs(bb, q, s.world_id, qw)
    assert np.array_equal(got, exp)
    assert (s.world_id[got[:, 1]] == qw[got[:, 0]]).all()
    assert np.array_equal(w.query_aabbs(q), O.query_aabbs(bb, q))      # NULL world = all worlds
    o = rng.uniform(-1, 7, (nq, 3))
    d = rng.normal(size=(nq, 3))
    rays = O.make_rays(o, d, 20.0)
    rays["world"] = qw
    gr = w.raycast(rays)
    er = O.raycast(s.xf, s.shapes, bb, rays, world_id=s.world_id, hull=s.hull)
    assert np.array_equal(gr["body"], er["body"])
    assert all(_bits_equal(gr[f], er[f]) for f in ("t", "nx", "ny", "nz", "flags"))
    hit = gr["body"] != axcd.NO_HIT
    assert hit.any() and (s.world_id[gr["body"][hit]] == qw[hit]).all()
    w.close()


def test_queries_tiny_scenes_and_errors():
    for n in (1, 2, 3):
        s = axcd.generate_scene(n, 3, 1.5)
        w = axcd.CollisionWorld.for_scene(s)
        with pytest.raises(axcd.AxcdError) as e:
            w.raycast(O.make_rays([[0, 0, 0]], [[1, 0, 0]], 1.0))    # no broadphase yet
        assert e.value.code == 503
        w.update()
        bb = w.aabbs()
        q = np.array([[-5, -5, -5, 5, 5, 5], [10, 10, 10, 11, 11, 11]], np.float32)
        assert np.array_equal(w.query_aabbs(q), O.query_aabbs(bb, q))
        rays = O.make_rays(s.xf[:, :3] - [0, 0, 4], [[0, 0, 1]] * n, 10.0)
        gr, er = w.raycast(rays), O.raycast(s.xf, s.shapes, bb, rays, hull=s.hull)
        assert gr.tobytes() == er.tobytes()
        assert (gr["body"] != axcd.NO_HIT).all()
        assert len(w.query_aabbs(np.zeros((0, 6), np.float32))) == 0
        w.close()


# ------------------------------------------------------------------ temporal coherence (rank 3) --
def _expand(tight, margin):
    m = np.float32(margin)
    return np.concatenate([tight[:, :3] - m, tight[:, 3:] + m], axis=1).astype(np.float32)


def test_temporal_coherence_fat_boxes_and_pair_cache():
    s = axcd.config_scene("C1", scale=0.2)
    margin = 0.05
    w = axcd.CollisionWorld.for_scene(s, aabbMargin=margin, flags=axcd.FLAG_TEMPORAL_COHERENCE, pairs_per_body=16)
    rng = np.random.default_rng(0)
    xf = s.xf.copy()
    fat = None
    skipped = []
    for step in range(7):
        if step in (1, 2, 4):                       # small motions: every tight box stays inside its fat box
            xf[:, :3] += rng.uniform(-0.012, 0.012, (s.n, 3)).astype(np.float32)
        if step == 3:                               # a few bodies jump: their fat boxes are rebuilt
            k = rng.choice(s.n, 50, replace=False)
            xf[k, :3] += rng.uniform(-1.5, 1.5, (50, 3)).astype(np.float32)
        w.set_transforms(xf)
        st = w.step()
        rc, tight = O.refit(xf, s.shapes, s.hull, nthreads=8)
        if fat is None:
            fat, moved = _expand(tight, margin), s.n
        else:
            inside = ((fat[:, :3] <= tight[:, :3]).all(1) & (tight[:, :3] <= tight[:, 3:]).all(1)
                      & (tight[:, 3:] <= fat[:, 3:]).all(1))
            fat[~inside] = _expand(tight[~inside], margin)
            moved = int((~inside).sum())
        assert st.movedBodies == moved, (step, st.movedBodies, moved)
        assert st.broadphaseSkipped == (1 if (moved == 0 and step > 0) else 0), step
        skipped.append(st.broadphaseSkipped)
        assert _bits_equal(w.aabbs(), fat)                                   # fat boxes, bit-exact
        pairs = O.broadphase(fat, nthreads=8)
        assert np.array_equal(w.pairs(), pairs)                             # candidate set of the fat boxes
        con, _, _ = O.narrowphase(xf, s.shapes, pairs, s.hull, nthreads=8)
        gc = w.contacts()
        assert gc.tobytes() == con.tobytes()
        # the contact set does not depend on the margin: same as the plain (tight-box) pipeline
        con_t, _, _ = O.narrowphase(xf, s.shapes, O.broadphase(tight, nthreads=8), s.hull, nthreads=8)
        assert np.array_equal(gc["a"], con_t["a"]) and np.array_equal(gc["b"], con_t["b"])
    assert skipped == [0, 1, 1, 0, 1, 1, 1], skipped
    # changing the candidate rule invalidates the cache even when nothing moved
    w.set_awake(np.ones(s.n, np.uint8))
    w.set_transforms(xf)
    assert w.step().broadphaseSkipped == 0
    w.close()
    # slab mode has no persistent fat-box state
    w = axcd.CollisionWorld(16, flags=axcd.FLAG_TEMPORAL_COHERENCE)
    with pytest.raises(axcd.AxcdError) as e:
        w.set_slab(0.0, 1.0)
    assert e.value.code == 600
    w.close()


def test_sleeping_pairs_are_dropped():
    s = axcd.config_scene("C1", scale=0.2)
    rng = np.random.default_rng(1)
    awake = (rng.random(s.n) < 0.5).astype(np.uint8)
    w = axcd.CollisionWorld.for_scene(s)
    w.set_awake(awake)
    st = w.step()
    rc, bb = O.refit(s.xf, s.shapes, s.hull, nthreads=8)
    allp = O.broadphase(bb, nthreads=8)
    keep = (awake[allp[:, 0]] | awake[allp[:, 1]]) != 0
    pairs = allp[keep]
    assert 0 < len(pairs) < len(allp)
    assert np.array_equal(w.pairs(), pairs)
    con, _, _ = O.narrowphase(s.xf, s.shapes, pairs, s.hull, nthreads=8)
    assert w.contacts().tobytes() == con.tobytes()
    w.set_awake(None)                              # rule off: the full candidate set again
    w.set_transforms(s.xf)
    w.step()
    assert np.array_equal(w.pairs(), allp)
    w.close()


# ------------------------------------------------------------------ GJK-based CCD (rank 4) ---------
def test_ccd_pairs_match_oracle():
    s, L = _query_scene(n=20000, seed=31, L=40.0)           # sparse: most pairs start separated
    rng = np.random.default_rng(6)
    w = axcd.CollisionWorld.for_scene(s)
    perm = rng.permutation(s.n).astype(np.uint32)             # every body in exactly one pair
    npairs = s.n // 2
    a, b = perm[:npairs], perm[npairs:]
    pairs = np.stack([a, b], axis=1).astype(np.uint32)
    # aim b at a with scatter, so that roughly half the sweeps hit
    aim = s.xf[a, :3] - s.xf[b, :3]
    disp = rng.normal(size=(s.n, 3)).astype(np.float32) * 0.05
    disp[b] = (aim * rng.uniform(0.8, 1.6, (npairs, 1)) + rng.normal(size=(npairs, 3)) * 0.3).astype(np.float32)
    got = w.ccd_pairs(pairs, disp)
    exp = O.ccd_pairs(s.xf, s.shapes, pairs, disp, s.hull)
    assert np.array_equal(got["hit"], exp["hit"])                      # bit-exact decisions
    assert np.array_equal(got["iterations"], exp["iterations"])
    for f in ("toi", "nx", "ny", "nz"):
        np.testing.assert_allclose(got[f], exp[f], rtol=REL, atol=REL * 1e-2)
    assert got.tobytes() == exp.tobytes()
    assert 0.15 < got["hit"].mean() < 0.9
    with pytest.raises(axcd.AxcdError) as e:
        w.ccd_pairs(np.array([[0, s.n]], np.uint32), disp)             # body index out of range
    assert e.value.code == 601
    w.close()


def test_ccd_with_rotation_matches_oracle():
    """axcd_ccd_pairs_angular: translation + first-order turning, bit for bit against the oracle."""
    s, _ = _query_scene(n=6000, seed=44, L=40.0)
    w = axcd.CollisionWorld.for_scene(s)
    rng = np.random.default_rng(15)
    perm = rng.permutation(s.n).astype(np.uint32)             # every body in exactly one pair
    a, b = perm[:s.n // 2], perm[s.n // 2:]
    pairs = np.stack([a, b], axis=1).astype(np.uint32)
    xf = s.xf.copy()
    xf[b, :3] = xf[a, :3] + (rng.normal(size=(len(b), 3)) * 2.0).astype(np.float32)
    w.set_transforms(xf)
    disp = (rng.normal(size=(s.n, 3)) * 0.05).astype(np.float32)
    disp[b] = ((xf[a, :3] - xf[b, :3]) * rng.uniform(0.3, 1.4, (len(b), 1)) + rng.normal(size=(len(b), 3)) * 0.3).astype(np.float32)
    rot = (rng.normal(size=(s.n, 3)) * rng.uniform(0.0, 1.2, (s.n, 1))).astype(np.float32)
    got = w.ccd_pairs(pairs, disp, rot)
    exp = O.ccd_pairs(xf, s.shapes, pairs, disp, s.hull, rot=rot)
    assert np.array_equal(got["hit"], exp["hit"])
    assert np.array_equal(got["iterations"], exp["iterations"])
    for f in ("toi", "nx", "ny", "nz"):
        np.testing.assert_allclose(got[f], exp[f], rtol=REL, atol=REL * 1e-2)
    assert got.tobytes() == exp.tobytes()
    assert 0.1 < got["hit"].mean() < 0.95
    lin = w.ccd_pairs(pairs, disp)
    assert (lin["toi"] != got["toi"]).mean() > 0.3                       # the turning matters
    with pytest.raises(axcd.AxcdError) as e:
        w.ccd_pairs(np.array([[0, s.n]], np.uint32), disp, rot)
    assert e.value.code == 601
    w.close()


def test_manifolds_capsule_box_mix():
    s, _ = _query_scene(n=20000, seed=41, L=24.0)     # boxes, spheres, capsules, hulls; dense enough to touch
    gm, bitwise = _manifold_parity(s, pairs_per_body=16)
    assert bitwise
    ta, tb = s.shapes["type"][gm["a"]], s.shapes["type"][gm["b"]]
    capbox = ((ta == 1) & (tb == 2)) | ((ta == 2) & (tb == 1))
    assert capbox.sum() > 200 and (gm["count"][capbox] == 2).sum() > 20
    capcap = (ta == 2) & (tb == 2)
    other = ~capbox & ~capcap & ~((ta == 1) & (tb == 1))
    assert (gm["count"][other] == 1).all()


@pytest.mark.gpu
def test_manifolds_log_pile_of_near_parallel_capsules():
    """Capsule-capsule manifolds: a dense pile of capsules whose axes are within a few degrees of a common
    direction (two thirds of them) or random (the rest) — the near-parallel pairs get two points."""
    rng = np.random.default_rng(51)
    s = axcd.generate_scene(30000, 51, 22.0, frac_box=1.0, frac_sphere=0.0)
    s.shapes["type"][:] = 2
    s.shapes["p0"][:] = rng.uniform(0.15, 0.3, s.n).astype(np.float32)
    s.shapes["p1"][:] = rng.uniform(0.8, 2.0, s.n).astype(np.float32)
    s.shapes["p2"][:] = 0.0
    common = O.axis_angle((1.0, 0.2, -0.4), 0.9)
    snap = rng.random(s.n) < 0.67
    for i in np.nonzero(snap)[0]:
        tilt = O.axis_angle(rng.normal(size=3), np.radians(rng.uniform(0.0, 7.0)))
        s.xf[i, 3:7] = O.quat_mul(tilt, common)
    s.xf[:, 7:10] = rng.uniform(0.8, 1.25, (s.n, 3)).astype(np.float32)
    gm, bitwise = _manifold_parity(s, pairs_per_body=24)
    assert bitwise
    assert (s.shapes["type"][gm["a"]] == 2).all()
    assert (gm["count"] == 2).sum() > 500 and (gm["count"] == 1).sum() > 500 and gm["count"].max() == 2


def test_temporal_coherence_two_refits_before_a_broadphase():
    """refit(); refit(); broadphase(): the second refit counts no moved body (the first one already rebuilt the
    fat boxes), so the cached candidate list must NOT be reused — the bodies moved by the first refit are not
    in it."""
    s = axcd.config_scene("C0")
    w = axcd.CollisionWorld.for_scene(s, aabbMargin=0.05, flags=axcd.FLAG_TEMPORAL_COHERENCE, pairs_per_body=16)
    w.step()
    xf = s.xf.copy()
    xf[:, :3] += np.float32(0.4) * np.sign(np.float32(5.0) - xf[:, :3])      # everybody moves out of its fat box, inwards
    w.set_transforms(xf)
    w.refit()
    w.aabbs()                 # e.g. a debug draw between the two
    w.refit()
    w._check(w._lib.axcd_broadphase(w._ctx), "axcd_broadphase")
    w.detect_collisions()
    st = w.stats()
    assert st.broadphaseSkipped == 0
    ref = axcd.CollisionWorld.for_scene(axcd.Scene(xf, s.shapes, s.hull), aabbMargin=0.05, pairs_per_body=16)
    ref.step()
    assert w.contacts().tobytes() == ref.contacts().tobytes()
    w.close()
    ref.close()
