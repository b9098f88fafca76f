"""One profiled step of the headline scene for `ncu --profile-from-start off`: two warm steps, then
cudaProfilerStart / refit + broadphase + narrowphase + manifolds + a batch of scene queries /
cudaProfilerStop.  Usage (on the GPU box):
    ncu --set full --clock-control none --import-source on --profile-from-start off \
        -o gpurun_out/step python profiles/one_step.py [workload]"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "axiom-physics-engine_b200"))
import axcd  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "headline"
s = axcd.config_scene(name)
w = axcd.CollisionWorld.for_scene(s)
for _ in range(2):
    w.step()
rng = np.random.default_rng(0)
nq = 1 << 16
L = float(s.xf[:, :3].max())
o = rng.uniform(0, L, (nq, 3)).astype(np.float32)
d = rng.normal(size=(nq, 3)).astype(np.float32)
d /= np.linalg.norm(d, axis=1, keepdims=True)
rays = np.zeros(nq, axcd.RAY_DT)
rays["ox"], rays["oy"], rays["oz"] = o[:, 0], o[:, 1], o[:, 2]
rays["dx"], rays["dy"], rays["dz"] = d[:, 0], d[:, 1], d[:, 2]
rays["tMax"] = 50.0
boxes = np.concatenate([o - 1.0, o + 1.0], axis=1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
flush.zero_()
torch.cuda.synchronize()
torch.cuda.profiler.start()
st = w.step()
w.build_manifolds()
hits = w.raycast(rays)
qh = w.query_aabbs(boxes)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(st.numPairs, st.numContacts, w.stats().contactPointCount, int((hits["body"] != axcd.NO_HIT).sum()), len(qh))
