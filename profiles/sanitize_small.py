"""Small end-to-end exercise of every kernel (step, manifolds, scene queries, CCD, temporal coherence,
sleeping, batched worlds, slab ghosts) for compute-sanitizer runs:
    compute-sanitizer --tool memcheck|racecheck|initcheck|synccheck python profiles/sanitize_small.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "axiom-physics-engine_b200"))
import axcd  # noqa: E402

rng = np.random.default_rng(0)
for name, kw in (("C0", {}), ("C2", {"scale": 0.003}), ("C3", {"scale": 8 / 4096})):
    s = axcd.config_scene(name, **kw)
    w = axcd.CollisionWorld.for_scene(s, pairs_per_body=16)
    st = w.step()
    w.build_manifolds()
    m, pts = w.manifolds()
    o = rng.uniform(-1, 11, (257, 3)).astype(np.float32)
    d = rng.normal(size=(257, 3)).astype(np.float32)
    rays = np.zeros(257, axcd.RAY_DT)
    rays["ox"], rays["oy"], rays["oz"] = o[:, 0], o[:, 1], o[:, 2]
    rays["dx"], rays["dy"], rays["dz"] = d[:, 0], d[:, 1], d[:, 2]
    rays["tMax"] = 20.0
    hits = w.raycast(rays)
    qh = w.query_aabbs(np.concatenate([o - 0.7, o + 0.7], axis=1))
    perm = rng.permutation(s.n).astype(np.uint32)
    pairs = np.stack([perm[: s.n // 2], perm[s.n // 2: 2 * (s.n // 2)]], axis=1)
    sw = w.ccd_pairs(pairs, rng.normal(size=(s.n, 3)).astype(np.float32))
    sw2 = w.ccd_pairs(pairs, rng.normal(size=(s.n, 3)).astype(np.float32), rng.normal(size=(s.n, 3)).astype(np.float32) * 0.5)
    # pose-only upload and the contact sink (fused narrowphase instantiation / copy kernel)
    sink = np.zeros(w.cfg.maxContacts, axcd.CONTACT_DT)
    axcd.pin_host_buffer(sink)
    w.set_contact_sink(sink.ctypes.data, w.cfg.maxContacts)
    w.set_poses(s.xf[:, :7])
    st3 = w.step()
    assert st3.numContacts == st.numContacts and sink[:st3.numContacts].tobytes() == w.contacts().tobytes()
    w.set_contact_sink(None, 0)
    axcd.unpin_host_buffer(sink)
    assert int(sw2["hit"].sum()) >= 0
    w.set_awake((rng.random(s.n) < 0.5).astype(np.uint8))
    w.set_transforms(s.xf)
    st2 = w.step()
    print(name, st.numPairs, st.numContacts, pts, int((hits["body"] != axcd.NO_HIT).sum()), len(qh), int(sw["hit"].sum()),
          st2.numPairs)
    w.close()
s = axcd.config_scene("C0")
w = axcd.CollisionWorld.for_scene(s, aabbMargin=0.05, flags=axcd.FLAG_TEMPORAL_COHERENCE, pairs_per_body=16)
a = w.step()
w.set_transforms(s.xf)
b = w.step()
print("coherent", a.numPairs, b.broadphaseSkipped, b.numContacts)
w.close()
