"""Generates tests/golden/c0_golden.npz: the C0 scene blob (1,000 boxes+spheres, L=10,
seed 1) plus a small hull-mix scene, with the CPU oracle's outputs for both.

    python tests/golden/make_golden.py

The reference snapshot has no collision code and no golden vectors, so these
fixtures pin the IN-REPO oracle (and the scene generator) against drift: compiler flags, libm,
refactors.  The GPU parity test compares the CUDA path against the same stored outputs."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "axiom-physics-engine_b200"))
import axcd  # noqa: E402
import oracle_lib as O  # noqa: E402


def run(scene):
    rc, bb = O.refit(scene.xf, scene.shapes, scene.hull)
    assert rc == 0
    pairs = O.broadphase(bb, brute=True)
    con, dist, _ = O.narrowphase(scene.xf, scene.shapes, pairs, scene.hull, want_distances=True)
    con2, _, _ = O.narrowphase(scene.xf, scene.shapes, pairs, scene.hull)
    assert np.array_equal(con, con2)
    # box-box through GJK/EPA instead of the closed-form SAT (AXCD_FLAG_BOXBOX_GJK_EPA)
    cong, _, _ = O.narrowphase(scene.xf, scene.shapes, pairs, scene.hull, cfg=O.default_cfg(False, True))
    return bb, pairs, con, dist, cong


out = {}
for tag, scene in (("c0", axcd.config_scene("C0")), ("c2s", axcd.config_scene("C2", scale=0.0005))):
    bb, pairs, con, dist, cong = run(scene)
    out.update({f"{tag}_xf": scene.xf, f"{tag}_shapes": scene.shapes, f"{tag}_hull": scene.hull,
                f"{tag}_aabb": bb, f"{tag}_pairs": pairs, f"{tag}_contacts": con, f"{tag}_dist": dist,
                f"{tag}_contacts_generic": cong})
    print(tag, scene.n, "bodies", len(pairs), "pairs", len(con), "contacts", len(cong), "contacts (generic box-box)")
np.savez_compressed(os.path.join(HERE, "c0_golden.npz"), **out)
