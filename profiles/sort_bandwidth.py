"""Device-resident bandwidth of the hand-written onesweep radix sort (key+value, 32-bit keys).

python profiles/sort_bandwidth.py  -> one JSON line per size.  Algorithmic bytes per element:
(8 B read + 8 B write) per pass + 4 B for the histogram read."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "axiom-physics-engine_b200"))
import axcd  # noqa: E402

peak = 6548.2
try:
    peak = float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"])
except Exception:
    pass
for n in (1 << 20, 1 << 22, 1 << 24, 1 << 26):
    w = axcd.CollisionWorld(n, max_pairs=1024)
    for bits in (24, 32):
        passes = (bits + 7) // 8
        nbytes = n * (16 * passes + 4)
        for method, ms in (("lsd onesweep (key, value)", w.sort_bench(n, bits, iters=10)),
                           ("lsd onesweep (key, identity payload)", w.sort_bench_morton(n, bits, iters=10, mode=0)),
                           ("bucket sort (key, identity payload)", w.sort_bench_morton(n, bits, iters=10, mode=1))):
            gbs = nbytes / (ms * 1e-3) / 1e9
            print(json.dumps({"n": n, "key_bits": bits, "method": method, "lsd_passes": passes, "ms": round(ms, 4),
                              "algorithmic_GB_of_the_lsd_sort": round(nbytes / 1e9, 4), "GBps": round(gbs, 1),
                              "frac_of_measured_hbm_peak": round(gbs / peak, 3)}))
    w.close()
