"""Where does the first axcd_query_aabbs call spend its time?  (1042 ms on the driver's box.)"""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "axiom-physics-engine_b200"))
import axcd
s = axcd.config_scene("headline")
w = axcd.CollisionWorld.for_scene(s)
w.step()
rng = np.random.default_rng(0)
nq = 1 << 18
o = rng.uniform(0, 100, (nq, 3)).astype(np.float32)
boxes = np.concatenate([o - 1.0, o + 1.0], axis=1)
for k in range(4):
    t0 = time.perf_counter()
    h = w.query_aabbs(boxes)
    print("call", k, round(1e3 * (time.perf_counter() - t0), 2), "ms", len(h), "hits", flush=True)
import ctypes as C
out = np.zeros((len(h) + 16, 2), np.uint32)
cnt = C.c_uint32(0)
for k in range(3):
    t0 = time.perf_counter()
    rc = w._lib.axcd_query_aabbs(w._ctx, boxes.ctypes.data_as(C.c_void_p), None, nq, out.ctypes.data_as(C.c_void_p), len(out), C.byref(cnt))
    print("raw C call", k, rc, round(1e3 * (time.perf_counter() - t0), 2), "ms", flush=True)
