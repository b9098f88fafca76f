"""Distribution of the polytope size EPA finishes with (vertices, face-slot high-water mark), from
the diagnostic oracle build (`make -C oracle hist`).  Used to size the CUDA fast-path caps.
usage: python profiles/epa_polytope_hist.py [C1|C2|headline] [scale]"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
sys.path.insert(0, os.path.join(ROOT, "axiom-physics-engine_b200"))
os.environ["AXREF_LIB"] = os.path.join(ROOT, "oracle", "_hist", "libaxref_hist.so")
import oracle_lib as O  # noqa: E402
import axcd  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "C1"
scale = float(sys.argv[2]) if len(sys.argv) > 2 else 1.0
sc = axcd.config_scene(name, scale)
nt = os.cpu_count()
_, aabb = O.refit(sc.xf, sc.shapes, sc.hull, nthreads=nt)
pairs = O.broadphase(aabb, nthreads=nt)
O.narrowphase(sc.xf, sc.shapes, pairs, sc.hull, nthreads=nt)
hv = (C.c_uint64 * 41)()
hf = (C.c_uint64 * 65)()
O.lib().axref_epa_hist(hv, hf)
hv = np.array(hv[:], dtype=np.int64)
hf = np.array(hf[:], dtype=np.int64)
tot = hv.sum()
print(f"{name}: {len(pairs)} pairs, {tot} EPA runs")
print("verts  count   cum%")
for i in np.nonzero(hv)[0]:
    print(f"{i:5d} {hv[i]:8d} {100.0 * hv[:i + 1].sum() / tot:6.2f}")
print("faces  count   cum%")
for i in np.nonzero(hf)[0]:
    print(f"{i:5d} {hf[i]:8d} {100.0 * hf[:i + 1].sum() / tot:6.2f}")
