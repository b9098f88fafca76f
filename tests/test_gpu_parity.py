"""GPU parity tests (run on the B200 box): the CUDA path through the C ABI vs the CPU oracle.

Integer / index results (candidate pairs, contact pair set, statuses) must be BIT-EXACT after the
canonical (a<b, sorted) ordering.  AABBs are float but compared bit-exactly too, because the
candidate set is defined on them by exact compares.  Contact floats (position, normal, depth) and
GJK distances: tolerance 1e-4 relative (BASELINE.json north_star), written as REL below.
"""
import ctypes as C

import numpy as np
import pytest

import axcd
import oracle_lib as O

pytestmark = pytest.mark.gpu
REL = 1e-4


def oracle_step(s, margin=0.0, brute=False, nthreads=8, want_distances=False, boxbox_generic=False):
    rc, bb = O.refit(s.xf, s.shapes, s.hull, margin=margin, nthreads=nthreads)
    assert rc == 0
    pairs = O.broadphase(bb, s.world_id, brute=brute, nthreads=nthreads)
    con, dist, st = O.narrowphase(s.xf, s.shapes, pairs, s.hull, nthreads=nthreads,
                                  cfg=O.default_cfg(want_distances, boxbox_generic), want_distances=want_distances)
    return bb, pairs, con, dist, st


def assert_bits_equal(a, b):
    """Bit-exact float comparison; NaNs must be in the same places (payload bits are not part of
    IEEE arithmetic and differ between x86 and the GPU)."""
    a, b = np.asarray(a, np.float32), np.asarray(b, np.float32)
    assert a.shape == b.shape
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb)
    assert np.array_equal(a.view(np.uint32)[~na], b.view(np.uint32)[~nb])


def assert_contacts_match(gc, oc):
    assert len(gc) == len(oc)
    assert np.array_equal(gc["a"], oc["a"]) and np.array_equal(gc["b"], oc["b"])   # bit-exact set + order
    assert np.array_equal(gc["status"], oc["status"])
    for f in ("px", "py", "pz", "nx", "ny", "nz", "depth"):
        np.testing.assert_allclose(gc[f], oc[f], rtol=REL, atol=REL * 1e-2)


def run_and_compare(s, margin=0.0, brute=False, boxbox_generic=False, nthreads=8, **kw):
    """boxbox_generic: box-box pairs through GJK/EPA on both sides (AXCD_FLAG_BOXBOX_GJK_EPA) instead of the
    closed-form SAT."""
    if boxbox_generic:
        kw["flags"] = kw.get("flags", 0) | axcd.FLAG_BOXBOX_GJK_EPA
    w = axcd.CollisionWorld.for_scene(s, aabbMargin=margin, **kw)
    st = w.step()
    bb, pairs, con, _, ost = oracle_step(s, margin=margin, brute=brute, boxbox_generic=boxbox_generic, nthreads=nthreads)
    assert_bits_equal(w.aabbs(), bb)                                              # bit-exact AABBs
    gp = w.pairs()
    assert st.numPairs == len(pairs)
    assert np.array_equal(gp, pairs)                                              # bit-exact pair set
    gc = w.contacts()
    assert st.numContacts == len(con)
    assert_contacts_match(gc, con)
    assert st.numPenetrating == ost.numPenetrating
    assert st.gjkFailures == ost.gjkFailures and st.epaFailures == ost.epaFailures
    bitwise = all(np.array_equal(gc[f].view(np.uint32), con[f].view(np.uint32))
                  for f in ("px", "py", "pz", "nx", "ny", "nz", "depth"))
    w.close()
    return st, bitwise


# ------------------------------------------------------------------ device primitives ----------
@pytest.mark.parametrize("n", [1, 2, 31, 4095, 4096, 4097, 100_000, 1_000_003])
def test_radix_sort_pairs32_vs_numpy(n):
    rng = np.random.default_rng(n)
    w = axcd.CollisionWorld(max(n, 16))
    for bits in (8, 24, 32):
        keys = rng.integers(0, 1 << bits, n, dtype=np.uint64).astype(np.uint32)
        if n > 10:
            keys[: n // 3] = keys[0]            # long runs of duplicates: stability matters
        vals = np.arange(n, dtype=np.uint32)
        k2, v2 = w.test_sort_pairs32(keys, vals, bits)
        order = np.argsort(keys, kind="stable")
        assert np.array_equal(k2, keys[order])
        assert np.array_equal(v2, vals[order])   # stable
    w.close()


@pytest.mark.parametrize("n", [1, 4097, 300_000])
def test_radix_sort_keys64_vs_numpy(n):
    rng = np.random.default_rng(n)
    w = axcd.CollisionWorld(16, max_pairs=max(n, 1024))
    for bits in (34, 40, 48):
        keys = rng.integers(0, 1 << bits, n, dtype=np.uint64)
        assert np.array_equal(w.test_sort_keys64(keys, bits), np.sort(keys))
    w.close()


# ------------------------------------------------------------------ parity per config ----------
BOXBOX_MODES = pytest.mark.parametrize("generic", [False, True], ids=["sat", "gjk_epa"])


@BOXBOX_MODES
def test_c0_reference_scene_vs_bruteforce_oracle(generic):
    st, bitwise = run_and_compare(axcd.config_scene("C0"), brute=True, boxbox_generic=generic)
    assert st.numPairs == 2875 and st.numContacts == 1381
    assert (st.numPenetrating > 0) == generic     # boxes and spheres only: EPA runs only when box-box is forced through it
    assert bitwise, "contact floats expected bit-identical (no FMA contraction on either side)"


@BOXBOX_MODES
def test_c1_100k_single_scene(generic):
    st, bitwise = run_and_compare(axcd.config_scene("C1"), boxbox_generic=generic)
    assert st.numPairs > 250_000
    assert bitwise


def test_traversal_heavy_tailed_sizes():
    # a few huge boxes among many small ones (deep, unbalanced overlap structure)
    rng = np.random.default_rng(7)
    n = 20000
    pos = rng.uniform(0, 30, (n, 3)).astype(np.float32)
    h = (rng.uniform(0.05, 1.0, (n, 3)) ** 4 * 8 + 0.05).astype(np.float32)
    xf = np.zeros((n, 10), np.float32)
    xf[:, :3] = pos
    xf[:, 6] = 1.0
    xf[:, 7:] = 1.0
    shapes = np.zeros(n, axcd.SHAPE_DT)
    shapes["type"] = 1
    shapes["p0"], shapes["p1"], shapes["p2"] = h[:, 0], h[:, 1], h[:, 2]
    s = axcd.Scene(xf, shapes)
    w = axcd.CollisionWorld.for_scene(s, pairs_per_body=400)
    st = w.step()
    rc, bb = O.refit(s.xf, s.shapes)
    pairs = O.broadphase(bb, nthreads=8)
    assert st.numPairs == len(pairs)
    assert np.array_equal(w.pairs(), pairs)
    w.close()


@BOXBOX_MODES
def test_c2_hull_mix_epa_heavy_scaled(generic):
    st, bitwise = run_and_compare(axcd.config_scene("C2", scale=0.05), boxbox_generic=generic)
    assert st.numPenetrating > (0.2 if generic else 0.1) * st.numPairs
    assert bitwise


def test_c2_hull_mix_full_size_1m():
    """Config C2 at its named size: 1 M bodies, 40 % boxes / 30 % spheres / 30 % 16-vertex hulls."""
    st, bitwise = run_and_compare(axcd.config_scene("C2"), nthreads=16)
    assert st.numBodies == 1_000_000 and st.numPenetrating > 0.1 * st.numPairs
    assert bitwise


def test_c3_all_4096_worlds_full_size():
    """Config C3 at its named size: 4096 independent 256-body worlds in one launch."""
    s = axcd.config_scene("C3")
    st, bitwise = run_and_compare(s, nthreads=16)
    assert st.numBodies == 4096 * 256 and bitwise


def test_c3_batched_worlds_no_cross_world_pairs():
    s = axcd.config_scene("C3", scale=64 / 4096)
    st, _ = run_and_compare(s)
    w = axcd.CollisionWorld.for_scene(s)
    w.step()
    p = w.pairs()
    assert (s.world_id[p[:, 0]] == s.world_id[p[:, 1]]).all()
    w.close()


def test_non_unit_scales_boxes_and_hulls():
    """Transform.scale enters refit (transformPoint) and the GJK cores (half-extents, hull vertices):
    non-uniform positive scales on a hull/box/sphere mix must stay bit-exact too."""
    s = axcd.config_scene("C2", scale=0.02)
    rng = np.random.default_rng(11)
    s.xf[:, 7:10] = rng.uniform(0.5, 1.6, (s.n, 3)).astype(np.float32)
    st, bitwise = run_and_compare(s)
    assert bitwise and st.numPenetrating > 0


def test_large_coordinates_far_from_origin():
    """Bodies 10^4 units from the origin: the A-centred GJK frame keeps the narrowphase well
    conditioned; sets stay bit-exact."""
    s = axcd.config_scene("C0")
    s.xf[:, :3] += np.float32(10000.0)
    st, bitwise = run_and_compare(s, brute=True)
    assert bitwise and st.numPairs > 2000


def test_aabb_margin_inflates_candidates():
    s = axcd.config_scene("C0")
    st0, _ = run_and_compare(s, brute=True)
    st1, _ = run_and_compare(s, margin=0.1, brute=True)
    assert st1.numPairs > st0.numPairs and st1.numContacts == st0.numContacts


@BOXBOX_MODES
def test_pair_distances_mode(generic):
    s = axcd.config_scene("C2", scale=0.01)
    w = axcd.CollisionWorld.for_scene(s, flags=axcd.FLAG_PAIR_DISTANCES | (axcd.FLAG_BOXBOX_GJK_EPA if generic else 0))
    w.step()
    _, pairs, con, dist, _ = oracle_step(s, want_distances=True, boxbox_generic=generic)
    assert np.array_equal(w.pairs(), pairs)
    assert_contacts_match(w.contacts(), con)
    gd = w.pair_distances()
    np.testing.assert_allclose(gd, dist, rtol=REL, atol=1e-6)
    w.close()


# ------------------------------------------------------------------ full size ------------------
@BOXBOX_MODES
def test_headline_1m_bodies_full_size(generic):
    s = axcd.config_scene("headline")
    w = axcd.CollisionWorld.for_scene(s, flags=axcd.FLAG_BOXBOX_GJK_EPA if generic else 0)
    st = w.step()
    # oracle grid broadphase and narrowphase finish in seconds on the box's host cores
    bb, pairs, con, _, ost = oracle_step(s, nthreads=16, boxbox_generic=generic)
    assert_bits_equal(w.aabbs(), bb)
    gp = w.pairs()
    assert np.array_equal(gp, pairs)
    key = gp[:, 0].astype(np.uint64) << np.uint64(32) | gp[:, 1]
    assert (gp[:, 0] < gp[:, 1]).all() and (np.diff(key.astype(np.int64)) > 0).all()   # canonical
    assert_contacts_match(w.contacts(), con)
    # idempotence: a second step on the same poses gives the same answer
    st2 = w.step()
    assert (st2.numPairs, st2.numContacts) == (st.numPairs, st.numContacts)
    assert np.array_equal(w.pairs(), gp)
    w.close()


# ------------------------------------------------------------------ edge cases -----------------
def test_empty_single_and_two_bodies():
    w = axcd.CollisionWorld(8)
    w.set_shapes(np.zeros(0, axcd.SHAPE_DT))
    w.set_transforms(np.zeros((0, 10), np.float32))
    st = w.step()
    assert (st.numPairs, st.numContacts) == (0, 0)
    w.set_shapes(np.array([O.sphere(1.0)], axcd.SHAPE_DT))
    w.set_transforms(np.array([O.xf()]))
    st = w.step()
    assert (st.numBodies, st.numPairs, st.numContacts) == (1, 0, 0)
    w.set_shapes(np.array([O.sphere(1.0), O.box(1, 1, 1)], axcd.SHAPE_DT))
    w.set_transforms(np.array([O.xf(), O.xf((1.5, 0, 0))]))
    st = w.step()
    assert (st.numPairs, st.numContacts) == (1, 1)
    c = w.contacts()[0]
    assert (c["a"], c["b"]) == (0, 1) and c["depth"] == pytest.approx(0.5, abs=1e-5)
    w.close()


def test_all_bodies_coincident_and_capacity_overflow():
    n = 200
    shapes = np.array([O.box(0.5, 0.5, 0.5)] * n, axcd.SHAPE_DT)
    xf = np.array([O.xf((1, 2, 3))] * n)
    s = axcd.Scene(xf, shapes)
    st, _ = run_and_compare(s, brute=True, pairs_per_body=100)
    assert st.numPairs == n * (n - 1) // 2
    w = axcd.CollisionWorld(n, max_pairs=1024)
    w.set_shapes(shapes)
    w.set_transforms(xf)
    with pytest.raises(axcd.AxcdError) as e:
        w.step()
    assert e.value.code == 601
    stats = axcd.Stats()
    assert w._lib.axcd_get_stats(w._ctx, C.byref(stats)) == 601
    assert stats.requiredPairs == n * (n - 1) // 2 and stats.numPairs == 1024
    w.close()


def test_nan_and_touching_boxes():
    shapes = np.array([O.box(0.5, 0.5, 0.5), O.box(0.5, 0.5, 0.5), O.sphere(0.5), O.sphere(0.5)], axcd.SHAPE_DT)
    xf = np.array([O.xf((0, 0, 0)), O.xf((1, 0, 0)), O.xf((np.nan, 0, 0)), O.xf((0.25, 0.25, 0))])
    st, _ = run_and_compare(axcd.Scene(xf, shapes), brute=True)
    w = axcd.CollisionWorld.for_scene(axcd.Scene(xf, shapes))
    w.step()
    p = w.pairs().tolist()
    assert [0, 1] in p                       # touching AABBs count (aabb.hpp:132-135)
    assert all(2 not in q for q in p)        # a NaN box has no pairs
    w.close()


def test_error_behaviour():
    w = axcd.CollisionWorld(8)
    with pytest.raises(axcd.AxcdError) as e:
        w.set_shapes(np.array([(3, 1.0, 1.0, 0.0)], axcd.SHAPE_DT))      # Plane
    assert e.value.code == 300
    with pytest.raises(axcd.AxcdError) as e:
        w.set_shapes(np.array([O.hull_shape(0, 4)], axcd.SHAPE_DT))       # hull range outside pool
    assert e.value.code == 300
    with pytest.raises(axcd.AxcdError) as e:
        w.set_shapes(np.zeros(9, axcd.SHAPE_DT))                          # over capacity
    assert e.value.code == 601
    assert w._lib.axcd_refit(w._ctx) == 503                               # no poses yet
    w.set_shapes(np.array([O.sphere(1.0)], axcd.SHAPE_DT))
    with pytest.raises(axcd.AxcdError) as e:
        w.set_transforms(np.zeros((2, 10), np.float32))                   # count mismatch
    assert e.value.code == 600
    w.set_transforms(np.array([O.xf()]))
    assert w._lib.axcd_narrowphase(w._ctx) == 503                         # broadphase not run
    # each stage runs once per refit (the per-step device counters are reset there)
    assert w._lib.axcd_refit(w._ctx) == 0
    assert w._lib.axcd_broadphase(w._ctx) == 0
    assert w._lib.axcd_broadphase(w._ctx) == 503
    assert w._lib.axcd_narrowphase(w._ctx) == 0
    assert w._lib.axcd_narrowphase(w._ctx) == 503
    assert w._lib.axcd_refit(w._ctx) == 0                                 # a new step may start at any point
    w.close()


def test_strided_transforms():
    s = axcd.config_scene("C0")
    w = axcd.CollisionWorld.for_scene(s)
    wide = np.zeros((s.n, 16), np.float32)
    wide[:, :10] = s.xf
    w.set_transforms(wide, stride=64)
    st = w.step()
    assert st.numPairs == 2875 and st.numContacts == 1381
    w.close()


def test_slab_sharding_with_cuda_backend_virtual_ranks():
    """The slab decomposition (axcd/sharding.py) driven by the CUDA backend, ranks emulated in
    sequence on one GPU: the union of the ranks' reports is the single-scene answer, bit for bit."""
    from axcd import sharding
    s = axcd.config_scene("C1", scale=0.2)
    be = sharding.CudaBackend()
    bb, ref_pairs, ref_con = be.step(s)
    cx = (bb[:, 0].astype(np.float64) + bb[:, 3]) * 0.5
    size = 4
    edges = sharding.plan_slabs(cx, size)
    owner = sharding.owner_of(cx, edges)
    got_p, got_c = [], []
    for r in range(size):
        mine = np.nonzero(owner == r)[0]
        owned = sharding._subset(s, mine)
        # ghosts that the other ranks would send to r
        ghosts = np.nonzero(sharding.slab_mask(bb, edges, r) & (owner != r))[0]
        local_idx = np.sort(np.concatenate([mine, ghosts]))
        local = sharding._subset(s, local_idx)
        lbb, lp, lc = be.step(local)
        gp, keep = sharding.filter_pairs_for_rank(lp, lbb, local_idx.astype(np.uint32), edges, r)
        got_p.append(gp)
        kept = {(int(a), int(b)) for a, b in lp[keep]}
        cm = np.array([(int(a), int(b)) in kept for a, b in zip(lc["a"], lc["b"])], bool)
        c = lc[cm].copy()
        c["a"], c["b"] = local_idx[c["a"]], local_idx[c["b"]]
        got_c.append(c)
        assert len(owned.xf) == len(mine)
    got = np.concatenate(got_p)
    key = got[:, 0].astype(np.uint64) << np.uint64(32) | got[:, 1]
    assert len(np.unique(key)) == len(key)
    assert np.array_equal(got[np.argsort(key)], ref_pairs)
    con = np.concatenate(got_c)
    con = con[np.lexsort((con["b"], con["a"]))]
    assert np.array_equal(con, ref_con)


def test_slab_rank_fast_path_virtual_ranks():
    """SlabRank: ownership rule + orientation by global id applied inside the traversal kernel
    (axcd_set_slab / axcd_set_body_keys / axcd_set_ghosts).  Ranks emulated in sequence on one GPU;
    the union of their reports must be the single-scene answer bit for bit."""
    from axcd import sharding
    s = axcd.config_scene("C1", scale=0.2)
    w = axcd.CollisionWorld.for_scene(s)
    w.step()
    bb, ref_pairs, ref_con = w.aabbs(), w.pairs().copy(), w.contacts().copy()
    w.close()
    size = 4
    cx = s.xf[:, 0]
    edges = sharding.plan_slabs(cx, size)
    owner = sharding.owner_of(cx, edges)
    ranks = []
    for r in range(size):
        mine = np.nonzero(owner == r)[0]
        ranks.append(sharding.SlabRank(sharding._subset(s, mine), mine.astype(np.uint32), edges, r, size))
    payloads = [rk.ghost_payloads(rk.owned_aabbs()) for rk in ranks]
    got_p, got_c = [], []
    for r, rk in enumerate(ranks):
        st = rk.step_with([payloads[o][r] for o in range(size) if o != r])
        assert st.numPairs > 0
        got_p.append(rk.pairs_global())
        got_c.append(rk.contacts_global())
        rk.close()
    got = np.concatenate(got_p)
    key = got[:, 0].astype(np.uint64) << np.uint64(32) | got[:, 1]
    assert len(np.unique(key)) == len(key), "a pair was reported by two ranks"
    assert np.array_equal(got[np.argsort(key)], ref_pairs)
    con = np.concatenate(got_c)
    con = con[np.lexsort((con["b"], con["a"]))]
    assert np.array_equal(con, ref_con)


def test_slab_rank_device_side_ghosts_match_host_side():
    """axcd_pack_ghosts / axcd_set_ghosts_device: ghost selection and hand-off entirely on the device
    must give every rank the same pair / contact counts as the host-side selection."""
    import torch
    from axcd import sharding
    s = axcd.config_scene("C1", scale=0.2)
    size = 3
    cx = s.xf[:, 0]
    edges = sharding.plan_slabs(cx, size)
    owner = sharding.owner_of(cx, edges)
    ranks = []
    for r in range(size):
        mine = np.nonzero(owner == r)[0]
        ranks.append(sharding.SlabRank(sharding._subset(s, mine), mine.astype(np.uint32), edges, r, size))
    host_payloads = [rk.ghost_payloads(rk.owned_aabbs()) for rk in ranks]
    host_counts = []
    for r, rk in enumerate(ranks):
        st = rk.step_with([host_payloads[o][r] for o in range(size) if o != r])
        host_counts.append((st.numBodies, st.numPairs, st.numContacts))
    dev_payloads = [rk.pack_device() for rk in ranks]
    for r, rk in enumerate(ranks):
        for o in range(size):
            if o != r:   # same ghost sets as the host path (order may differ)
                assert dev_payloads[o][r].numel() // 16 == (len(host_payloads[o][r]) - 2) // 15
        st = rk.step_with_device([dev_payloads[o][r].clone() for o in range(size) if o != r])
        assert (st.numBodies, st.numPairs, st.numContacts) == host_counts[r]
    total_pairs = sum(c[1] for c in host_counts)
    w = axcd.CollisionWorld.for_scene(s)
    assert w.step().numPairs == total_pairs
    w.close()
    for rk in ranks:
        rk.close()
    torch.cuda.synchronize()


def test_degenerate_pair_zoo_bit_exact():
    """Thousands of isolated pairs in awkward configurations — coincident centres, axis-aligned
    rotations, exactly touching faces, zero-size boxes, zero-height capsules, duplicate hull
    vertices, wildly different scales — all shape combinations.  Every pair sits in its own cell far
    from the others, so the candidate set is exactly the intended pairs; the CUDA path must agree
    with the oracle bit for bit on all of them."""
    rng = np.random.default_rng(1234)
    npairs = 6000
    hull_pool = []
    xf = np.zeros((2 * npairs, 10), np.float32)
    shapes = np.zeros(2 * npairs, axcd.SHAPE_DT)

    def rand_quat(k):
        mode = k % 4
        if mode == 0:
            return (0, 0, 0, 1)                                             # identity
        if mode == 1:
            return O.axis_angle(np.eye(3)[rng.integers(0, 3)], np.pi / 2 * rng.integers(0, 4))   # axis-aligned
        return O.axis_angle(rng.normal(size=3), rng.uniform(0, 6.28))

    def rand_shape(k):
        kind = rng.integers(0, 4)
        if kind == 0:
            return (0, [0.0, 1e-3, 0.3, 0.5][rng.integers(0, 4)], 0, 0)
        if kind == 1:
            h = rng.uniform(0.2, 0.5, 3)
            if k % 7 == 0:
                h[rng.integers(0, 3)] = 0.0                                  # flat box
            return (1, h[0], h[1], h[2])
        if kind == 2:
            return (2, rng.uniform(0.05, 0.3), [0.0, 0.6][rng.integers(0, 2)], 0)   # some zero-height capsules
        first = len(hull_pool)
        v = rng.normal(size=(8, 3)) * 0.3
        if k % 5 == 0:
            v[4:] = v[:4]                                                    # duplicate vertices
        hull_pool.extend(v.tolist())
        return (4, np.array([first], np.uint32).view(np.float32)[0], np.array([8], np.uint32).view(np.float32)[0], 0)

    for k in range(npairs):
        cell = np.array([k % 40, (k // 40) % 40, k // 1600], np.float32) * 10.0
        offs = [np.zeros(3), np.array([0.5, 0, 0]), np.array([1.0, 0, 0]), rng.uniform(-0.8, 0.8, 3)][k % 4]
        for side in range(2):
            i = 2 * k + side
            xf[i, :3] = cell + (offs if side else 0)
            xf[i, 3:7] = rand_quat(k + side)
            xf[i, 7:10] = [1, 1, 1] if k % 3 else rng.uniform(0.25, 3.0, 3)
            shapes[i] = rand_shape(k + side)
    s = axcd.Scene(xf, shapes, np.array(hull_pool, np.float32).reshape(-1, 3) if hull_pool else None)
    st, bitwise = run_and_compare(s)
    assert bitwise
    assert st.numPairs > npairs // 2 and st.numContacts > npairs // 4
    st, bitwise = run_and_compare(s, boxbox_generic=True)
    assert bitwise


def test_deep_tree_everything_overlaps():
    """Worst case for the traversal stacks: a deep LBVH (Morton keys with long common prefixes: positions
    in a geometric progression, plus exact duplicates that are split by the index tie-break) in which
    every box overlaps every other, so every subtree is entered.  A stack overflow would surface as
    error 505 (never as silently missing pairs); the pair set must be all n(n-1)/2 pairs."""
    n = 2400
    xf = np.zeros((n, 10), np.float32)
    k = np.arange(n)
    base = (100.0 * 2.0 ** -(k % 24)).astype(np.float32)          # 24 nested scales ...
    xf[:, 0], xf[:, 1], xf[:, 2] = base, base * 0.5, base * 0.25   # ... 100 exact duplicates of each
    xf[:, 6] = 1.0
    xf[:, 7:] = 1.0
    shapes = np.zeros(n, axcd.SHAPE_DT)
    shapes["type"] = 0
    shapes["p0"] = 500.0                                          # every sphere's box covers the scene
    s = axcd.Scene(xf, shapes)
    w = axcd.CollisionWorld(n, max_pairs=n * (n - 1) // 2 + 16, max_contacts=16)
    w.set_shapes(s.shapes)
    w.set_transforms(s.xf)
    w.update()
    st = w.stats()
    assert st.numPairs == n * (n - 1) // 2
    gp = w.pairs()
    iu = np.triu_indices(n, 1)
    assert np.array_equal(gp, np.stack(iu, axis=1).astype(np.uint32))
    # the scene queries walk the same tree with their own stacks
    hits = w.query_aabbs(np.array([[-1e4, -1e4, -1e4, 1e4, 1e4, 1e4]], np.float32))
    assert np.array_equal(hits[:, 1], np.arange(n, dtype=np.uint32))
    w.close()


# ------------------------------------------------------------------ the step as a CUDA graph ---
def test_graph_replay_matches_direct_launches():
    """axcd_step replays a CUDA graph of the whole step once the launch configuration has been stable for a
    step; results must be identical to the direct launches (AXCD_FLAG_NO_GRAPH), step after step, also when
    the poses change between steps."""
    s = axcd.config_scene("C1", scale=0.3)
    wg = axcd.CollisionWorld.for_scene(s)
    wd = axcd.CollisionWorld.for_scene(s, flags=axcd.FLAG_NO_GRAPH)
    rng = np.random.default_rng(5)
    seen_graph = False
    for it in range(6):
        xf = s.xf.copy()
        xf[:, :3] += rng.normal(size=(s.n, 3)).astype(np.float32) * 0.05 * it
        wg.set_transforms(xf)
        wd.set_transforms(xf)
        sg, sd = wg.step(), wd.step()
        seen_graph = seen_graph or bool(sg.graphLaunched)
        assert not sd.graphLaunched
        assert (sg.numPairs, sg.numContacts, sg.kernelLaunches) == (sd.numPairs, sd.numContacts, sd.kernelLaunches)
        assert np.array_equal(wg.pairs(), wd.pairs())
        assert wg.contacts().tobytes() == wd.contacts().tobytes()
        assert_bits_equal(wg.aabbs(), wd.aabbs())
    assert seen_graph, "the steady-state step must be a graph launch"
    # a configuration change (filters on) drops the graph; the next steps stay correct
    filt = np.stack([np.ones(s.n), np.full(s.n, 0xFFFF), rng.integers(-2, 3, s.n)], axis=1)
    wg.set_filters(filt)
    wd.set_filters(filt)
    for _ in range(3):
        sg, sd = wg.step(), wd.step()
        assert np.array_equal(wg.pairs(), wd.pairs()) and sg.numPairs < s.n * 8
    # staged calls after graph steps still give per-stage timings
    wg.update()
    wg.detect_collisions()
    st = wg.stats()
    assert st.pairMs > 0 and st.graphLaunched == 0
    wg.close()
    wd.close()


def test_filter_and_awake_argument_checks():
    s = axcd.config_scene("C0")
    w = axcd.CollisionWorld.for_scene(s)
    with pytest.raises(axcd.AxcdError) as e:
        w.set_filters(np.ones((s.n - 1, 3)))          # one record per body, no fewer
    assert e.value.code == 600
    with pytest.raises(axcd.AxcdError) as e:
        w.set_awake(np.ones(s.n - 1))
    assert e.value.code == 600
    with pytest.raises(axcd.AxcdError) as e:
        w.set_shapes(np.array([(1, -1.0, 1.0, 1.0)], axcd.SHAPE_DT))   # negative half extent
    assert e.value.code == 300
    with pytest.raises(axcd.AxcdError) as e:
        w.set_shapes(np.array([(0, np.inf, 0.0, 0.0)], axcd.SHAPE_DT))  # non-finite radius
    assert e.value.code == 300
    w.close()


def test_refit_mat4_route_matches_oracle_bit_for_bit():
    """AXCD_FLAG_REFIT_MAT4_ROUTE: boxes refit through AABB::transform(Transform::toMatrix())
    as src/math/aabb.cpp:8-35 does it.  AABBs bit-identical to the oracle's restatement of that route (which
    tests/test_oracle_vs_reference.py pins to the compiled reference functions), pair and contact sets exact."""
    s = axcd.config_scene("C2", scale=0.02)
    s.xf[:, 7:10] = np.random.default_rng(3).uniform(0.5, 1.6, (s.n, 3)).astype(np.float32)
    w = axcd.CollisionWorld.for_scene(s, flags=axcd.FLAG_REFIT_MAT4_ROUTE)
    st = w.step()
    rc, bb = O.refit(s.xf, s.shapes, s.hull, mat4_route=True)
    assert_bits_equal(w.aabbs(), bb)
    rc, bb0 = O.refit(s.xf, s.shapes, s.hull)
    assert not np.array_equal(bb.view(np.uint32), bb0.view(np.uint32))     # the two routes round differently
    pairs = O.broadphase(bb, nthreads=8)
    assert np.array_equal(w.pairs(), pairs)
    con, _, _ = O.narrowphase(s.xf, s.shapes, pairs, s.hull, nthreads=8)
    assert w.contacts().tobytes() == con.tobytes()
    w.close()


# ------------------------------------------------------------------ the step's Morton sort ------
@pytest.mark.parametrize("n", [4096, 100_000, 1_000_003])
@pytest.mark.parametrize("bits", [12, 24, 30])
def test_morton_bucket_sort_vs_numpy(n, bits):
    """The step sorts (Morton key, body index) with one MSD bucket pass + per-bucket shared-memory sorts, and falls
    back to the LSD radix kernels on the device when a bucket overflows.  Both must give the stable order."""
    rng = np.random.default_rng(n + bits)
    w = axcd.CollisionWorld(max(n, 16))
    uniform = rng.integers(0, 1 << bits, n, dtype=np.uint64).astype(np.uint32)
    clustered = (rng.integers(0, 1 << min(bits, 10), n, dtype=np.uint64)).astype(np.uint32)    # top bits all zero: one bucket
    few = np.repeat(rng.integers(0, 1 << bits, 7, dtype=np.uint64).astype(np.uint32), n // 7 + 1)[:n]   # 7 distinct keys
    for keys, expect_fb in ((uniform, 0), (clustered, 1 if bits >= 24 else None), (few, 1)):
        order = np.argsort(keys, kind="stable")
        for mode in (1, 0):
            k2, v2, fb = w.test_sort_morton(keys, bits, mode)
            assert np.array_equal(k2, keys[order]) and np.array_equal(v2, order.astype(np.uint32)), (mode, fb)
            if mode == 0:
                assert fb == 1
            elif expect_fb is not None and bits >= 12 and n >= 100_000:
                assert fb == expect_fb, (fb, expect_fb)
    w.close()


def test_clustered_scene_takes_the_sort_fallback_and_stays_exact():
    """Most bodies piled into one corner of a huge scene box: the Morton keys collapse into a few buckets, the
    bucket sort gives up on the device and the LSD kernels sort instead — same pairs, same contacts."""
    import os
    os.environ["AXCD_BUCKET_SORT"] = "1"      # read at axcd_create
    try:
        s = axcd.config_scene("C1", scale=0.5)
        s.xf[:, :3] *= np.float32(0.5)
        s.xf[0, :3] = 4000.0          # one far outlier stretches the scene box
        st, bitwise = run_and_compare(s, pairs_per_body=96)
        assert bitwise and st.sortFallback == 1 and st.sortMaxBucket > 1024
        st2, bitwise2 = run_and_compare(axcd.config_scene("C1", scale=0.5))
        assert bitwise2 and st2.sortFallback == 0 and 0 < st2.sortMaxBucket <= 1024
    finally:
        del os.environ["AXCD_BUCKET_SORT"]


def test_fused_closed_form_narrowphase_matches_the_split_kernels():
    """Sphere / box scenes run the whole narrowphase in one kernel (classify + closed forms + in-order compaction);
    AXCD_SPLIT_NARROW=1 keeps the three separate kernels.  Same contacts, byte for byte, same order."""
    import os
    s = axcd.config_scene("C1")
    w1 = axcd.CollisionWorld.for_scene(s)
    st1 = w1.step()
    os.environ["AXCD_SPLIT_NARROW"] = "1"
    try:
        w2 = axcd.CollisionWorld.for_scene(s)
    finally:
        del os.environ["AXCD_SPLIT_NARROW"]
    st2 = w2.step()
    assert st1.kernelLaunches < st2.kernelLaunches
    assert (st1.numPairs, st1.numContacts) == (st2.numPairs, st2.numContacts)
    assert w1.contacts().tobytes() == w2.contacts().tobytes()
    w1.close()
    w2.close()


@pytest.mark.gpu
def test_set_poses_uploads_position_and_rotation_only():
    """axcd_set_poses: 28 B per body, scales stay as the last set_transforms left them; packed and strided."""
    s = axcd.config_scene("C1", scale=0.2)
    rng = np.random.default_rng(8)
    s.xf[:, 7:10] = rng.uniform(0.7, 1.3, (s.n, 3)).astype(np.float32)
    w = axcd.CollisionWorld(s.n, max_pairs=24 * s.n, max_hull_verts=len(s.hull))
    w.set_shapes(s.shapes, s.hull, s.world_id)
    with pytest.raises(axcd.AxcdError) as e:
        w.set_poses(s.xf[:, :7])                      # no scales on the device yet
    assert e.value.code == 503
    w.set_transforms(s.xf)
    w.step()
    xf2 = s.xf.copy()
    xf2[:, :3] += rng.normal(size=(s.n, 3)).astype(np.float32) * 0.2
    q = xf2[:, 3:7] + rng.normal(size=(s.n, 4)).astype(np.float32) * 0.2
    xf2[:, 3:7] = (q / np.linalg.norm(q, axis=1, keepdims=True)).astype(np.float32)
    for mode in ("packed", "strided"):
        if mode == "packed":
            w.set_poses(xf2[:, :7])
        else:
            junk = xf2.copy()
            junk[:, 7:10] = 99.0                      # must not be read
            w.set_poses(junk, stride=40)
        st = w.step()
        rc, bb = O.refit(xf2, s.shapes, s.hull, nthreads=8)
        pairs = O.broadphase(bb, nthreads=8)
        con, _, _ = O.narrowphase(xf2, s.shapes, pairs, s.hull, nthreads=8)
        assert np.array_equal(w.aabbs(), bb)
        assert np.array_equal(w.pairs(), pairs)
        assert w.contacts().tobytes() == con.tobytes() and st.numContacts == len(con) > 0
        xf2[:, :3] += 0.01                            # a different pose for the second mode
    with pytest.raises(axcd.AxcdError) as e:
        w.set_poses(xf2[:, :7], stride=24)
    assert e.value.code == 600
    w.close()


@pytest.mark.gpu
@pytest.mark.parametrize("scene_name,kw", [("C1", {}), ("C2", {"scale": 0.05})], ids=["closed-form-fused", "generic-gjk-epa"])
def test_contact_sink_receives_the_steps_contacts(scene_name, kw):
    """axcd_set_contact_sink: the step delivers the contacts into a page-locked host buffer (copy-engine chunks that
    follow the fused narrowphase kernel tile by tile, or a copy kernel behind GJK / EPA); same records, same order."""
    s = axcd.config_scene(scene_name, **kw)
    w = axcd.CollisionWorld.for_scene(s, pairs_per_body=16)
    cap = w.cfg.maxContacts
    sink = np.zeros(cap, axcd.CONTACT_DT)
    with pytest.raises(axcd.AxcdError) as e:
        w.set_contact_sink(sink.ctypes.data, cap)         # pageable memory
    assert e.value.code == 600
    axcd.pin_host_buffer(sink)
    try:
        with pytest.raises(axcd.AxcdError) as e:
            w.set_contact_sink(sink.ctypes.data, cap - 1)  # too small
        assert e.value.code == 600
        w.set_contact_sink(sink.ctypes.data, cap)
        for rep in range(3):                               # direct launches, then the captured graph
            sink[:] = 0
            w.set_transforms(s.xf)
            st = w.step()
            ref = w.contacts()
            assert st.numContacts == len(ref) > 0
            assert sink[:len(ref)].tobytes() == ref.tobytes()
            assert not sink[len(ref):].view(np.uint32).any()
        con, _, _ = O.narrowphase(s.xf, s.shapes, w.pairs(), s.hull, nthreads=8)
        assert sink[:len(con)].tobytes() == con.tobytes()
        w.set_contact_sink(None, 0)                        # detached: the buffer is left alone
        sink[:] = 0
        w.set_transforms(s.xf)
        w.step()
        assert not sink.view(np.uint32).any()
    finally:
        w.close()
        axcd.unpin_host_buffer(sink)


def test_contact_sink_survives_steps_queued_back_to_back():
    """Two fused steps queued without a blocking call between them: the second launch first completes the first
    step's sink (its copy-engine chunks follow progress words the second launch must not clear early), and the sink
    ends up holding the second step's contacts.  Detaching the sink while a step is in flight completes it too."""
    s = axcd.config_scene("C1")
    w = axcd.CollisionWorld.for_scene(s, pairs_per_body=16)
    cap = w.cfg.maxContacts
    sink = np.zeros(cap, axcd.CONTACT_DT)
    axcd.pin_host_buffer(sink)
    try:
        w.set_contact_sink(sink.ctypes.data, cap)
        xf2 = s.xf.copy()
        xf2[:, 0] += np.float32(0.01) * (np.arange(s.n) % 7).astype(np.float32)   # a slightly different scene
        for rep in range(3):   # direct launches first, then graph replays
            w.set_transforms(s.xf)
            w.step_async()
            w.set_transforms(xf2)
            w.step_async()
            st = w.stats()
            ref = w.contacts()
            assert st.numContacts == len(ref) > 0
            assert sink[:len(ref)].tobytes() == ref.tobytes()
        con, _, _ = O.narrowphase(xf2, s.shapes, w.pairs(), s.hull, nthreads=8)
        assert sink[:len(con)].tobytes() == con.tobytes()
        w.set_transforms(s.xf)
        w.step_async()
        w.set_contact_sink(None, 0)          # detach with the step in flight: that step still delivers
        st = w.stats()
        ref = w.contacts()
        assert sink[:len(ref)].tobytes() == ref.tobytes()
    finally:
        w.close()
        axcd.unpin_host_buffer(sink)

