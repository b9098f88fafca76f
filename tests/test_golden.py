"""Committed golden fixtures (tests/golden/c0_golden.npz, made by tests/golden/make_golden.py).

They pin the in-repo oracle and the scene generator (the reference has no golden vectors for this
path); the GPU test checks the CUDA path against the same stored outputs."""
import os

import numpy as np
import pytest

import axcd
import oracle_lib as O

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "c0_golden.npz"))


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


@pytest.mark.parametrize("tag", ["c0", "c2s"])
def test_oracle_reproduces_golden_outputs(tag):
    xf, shapes, hull = G[f"{tag}_xf"], G[f"{tag}_shapes"], G[f"{tag}_hull"]
    rc, bb = O.refit(xf, shapes, hull)
    assert rc == 0 and np.array_equal(bits(bb), bits(G[f"{tag}_aabb"]))
    pairs = O.broadphase(bb)
    assert np.array_equal(pairs, G[f"{tag}_pairs"])
    con, dist, _ = O.narrowphase(xf, shapes, pairs, hull, want_distances=True)
    gold = G[f"{tag}_contacts"]
    assert np.array_equal(con["a"], gold["a"]) and np.array_equal(con["b"], gold["b"])
    for f in ("px", "py", "pz", "nx", "ny", "nz", "depth"):
        assert np.array_equal(bits(con[f]), bits(gold[f])), f
    assert np.array_equal(bits(dist), bits(G[f"{tag}_dist"]))
    # box-box through GJK/EPA instead of the closed-form SAT
    cong, _, _ = O.narrowphase(xf, shapes, pairs, hull, cfg=O.default_cfg(False, True))
    assert cong.tobytes() == G[f"{tag}_contacts_generic"].tobytes()


def test_scene_generator_reproduces_golden_blobs():
    s = axcd.config_scene("C0")
    assert np.array_equal(bits(s.xf), bits(G["c0_xf"])) and np.array_equal(s.shapes, G["c0_shapes"])
    s = axcd.config_scene("C2", scale=0.0005)
    assert np.array_equal(bits(s.xf), bits(G["c2s_xf"])) and np.array_equal(bits(s.hull), bits(G["c2s_hull"]))


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["c0", "c2s"])
def test_cuda_path_reproduces_golden_outputs(tag):
    s = axcd.Scene(G[f"{tag}_xf"], G[f"{tag}_shapes"], G[f"{tag}_hull"])
    w = axcd.CollisionWorld.for_scene(s, flags=axcd.FLAG_PAIR_DISTANCES)
    w.step()
    assert np.array_equal(bits(w.aabbs()), bits(G[f"{tag}_aabb"]))
    assert np.array_equal(w.pairs(), G[f"{tag}_pairs"])
    con, gold = w.contacts(), G[f"{tag}_contacts"]
    assert np.array_equal(con["a"], gold["a"]) and np.array_equal(con["b"], gold["b"])
    for f in ("px", "py", "pz", "nx", "ny", "nz", "depth"):
        np.testing.assert_allclose(con[f], gold[f], rtol=1e-4, atol=1e-6)     # stated FP32 tolerance
        assert np.array_equal(bits(con[f]), bits(gold[f])), f                 # and in fact bit-identical
    np.testing.assert_allclose(w.pair_distances(), G[f"{tag}_dist"], rtol=1e-4, atol=1e-6)
    w.close()
    # box-box through GJK/EPA (AXCD_FLAG_BOXBOX_GJK_EPA): the generic kernels on the same pairs
    w = axcd.CollisionWorld.for_scene(s, flags=axcd.FLAG_BOXBOX_GJK_EPA)
    w.step()
    assert w.contacts().tobytes() == G[f"{tag}_contacts_generic"].tobytes()
    w.close()


# ---- the stages on top of the hot path (tests/golden/make_golden_next.py) ---------------------------------
GN = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "next_rows_golden.npz"))


@pytest.mark.parametrize("tag", ["c0", "mix"])
def test_oracle_reproduces_next_row_goldens(tag):
    xf, shapes, hull = GN[f"{tag}_xf"], GN[f"{tag}_shapes"], GN[f"{tag}_hull"]
    rc, bb = O.refit(xf, shapes, hull)
    con, _, _ = O.narrowphase(xf, shapes, O.broadphase(bb), hull)
    assert con.tobytes() == GN[f"{tag}_contacts"].tobytes()
    man, pts = O.manifolds(xf, shapes, con)
    assert man.tobytes() == GN[f"{tag}_manifolds"].tobytes() and pts == int(GN[f"{tag}_points"])
    assert O.raycast(xf, shapes, bb, GN[f"{tag}_rays"], hull=hull).tobytes() == GN[f"{tag}_rayhits"].tobytes()
    assert np.array_equal(O.query_aabbs(bb, GN[f"{tag}_qboxes"]), GN[f"{tag}_qhits"])
    assert O.ccd_pairs(xf, shapes, GN[f"{tag}_cpairs"], GN[f"{tag}_disp"], hull).tobytes() == GN[f"{tag}_sweeps"].tobytes()


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["c0", "mix"])
def test_cuda_path_reproduces_next_row_goldens(tag):
    s = axcd.Scene(GN[f"{tag}_xf"], GN[f"{tag}_shapes"], GN[f"{tag}_hull"])
    w = axcd.CollisionWorld.for_scene(s, pairs_per_body=16)
    w.step()
    assert w.contacts().tobytes() == GN[f"{tag}_contacts"].tobytes()
    w.build_manifolds()
    man, pts = w.manifolds()
    assert man.tobytes() == GN[f"{tag}_manifolds"].tobytes() and pts == int(GN[f"{tag}_points"])
    assert w.raycast(GN[f"{tag}_rays"]).tobytes() == GN[f"{tag}_rayhits"].tobytes()
    assert np.array_equal(w.query_aabbs(GN[f"{tag}_qboxes"]), GN[f"{tag}_qhits"])
    assert w.ccd_pairs(GN[f"{tag}_cpairs"], GN[f"{tag}_disp"]).tobytes() == GN[f"{tag}_sweeps"].tobytes()
    w.close()
