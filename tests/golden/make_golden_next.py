"""Generates tests/golden/next_rows_golden.npz: the CPU oracle's outputs of the stages built on top of
the hot path (contact manifolds, scene queries, GJK-based CCD) on the C0 scene and a small
capsule/hull mix, with their inputs.

    python tests/golden/make_golden_next.py

Like c0_golden.npz these pin the IN-REPO oracle against drift (the reference has no vectors for any of
this); tests/test_golden.py replays them through the oracle (CPU) and through the CUDA path (GPU)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "axiom-physics-engine_b200"))
import axcd  # noqa: E402
import oracle_lib as O  # noqa: E402


def mix_scene():
    s = axcd.generate_scene(600, 17, 8.0, frac_box=0.4, frac_sphere=0.4)      # 20 % hulls
    k = np.where(s.shapes["type"] == 0)[0][::2]
    s.shapes["type"][k] = 2                                                     # capsules
    s.shapes["p0"][k] = np.float32(0.2)
    s.shapes["p1"][k] = np.float32(0.7)
    s.xf[:, 7:10] = np.random.default_rng(17).uniform(0.7, 1.4, (s.n, 3)).astype(np.float32)
    return s


out = {}
for tag, s, L in (("c0", axcd.config_scene("C0"), 10.0), ("mix", mix_scene(), 8.0)):
    rng = np.random.default_rng(5)
    rc, bb = O.refit(s.xf, s.shapes, s.hull)
    pairs = O.broadphase(bb, brute=True)
    con, _, _ = O.narrowphase(s.xf, s.shapes, pairs, s.hull)
    man, pts = O.manifolds(s.xf, s.shapes, con)
    o = rng.uniform(-1, L + 1, (300, 3)).astype(np.float32)
    d = rng.normal(size=(300, 3)).astype(np.float32)
    rays = O.make_rays(o, d, 25.0)
    hits = O.raycast(s.xf, s.shapes, bb, rays, hull=s.hull)
    qboxes = np.concatenate([o - 0.6, o + 0.6], axis=1).astype(np.float32)
    qhits = O.query_aabbs(bb, qboxes)
    perm = rng.permutation(s.n).astype(np.uint32)
    cpairs = np.stack([perm[: s.n // 2], perm[s.n // 2: 2 * (s.n // 2)]], axis=1)
    disp = rng.normal(size=(s.n, 3)).astype(np.float32) * 0.05
    disp[cpairs[:, 1]] = ((s.xf[cpairs[:, 0], :3] - s.xf[cpairs[:, 1], :3]) * rng.uniform(0.6, 1.4, (len(cpairs), 1))
                          + rng.normal(size=(len(cpairs), 3)) * 0.7).astype(np.float32)
    sweeps = O.ccd_pairs(s.xf, s.shapes, cpairs, disp, s.hull)
    out.update({f"{tag}_xf": s.xf, f"{tag}_shapes": s.shapes, f"{tag}_hull": s.hull, f"{tag}_contacts": con,
                f"{tag}_manifolds": man, f"{tag}_points": np.uint64(pts), f"{tag}_rays": rays, f"{tag}_rayhits": hits,
                f"{tag}_qboxes": qboxes, f"{tag}_qhits": qhits, f"{tag}_cpairs": cpairs, f"{tag}_disp": disp,
                f"{tag}_sweeps": sweeps})
    print(tag, s.n, "bodies", len(con), "contacts", pts, "points", int((hits["body"] != O.NO_HIT).sum()), "ray hits",
          len(qhits), "query hits", int(sweeps["hit"].sum()), "sweep hits")
np.savez_compressed(os.path.join(HERE, "next_rows_golden.npz"), **out)
