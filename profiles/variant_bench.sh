#!/bin/bash
# Benchmarks every variants/libaxcd_*.so (AXCD_LIB) on the headline scene: one line per variant.
tag=${1:-var}
wl=${2:-headline}
out=gpurun_out/$tag
mkdir -p $out
for so in base axiom-physics-engine_b200/variants/libaxcd_*.so; do
  name=$(basename $so .so); name=${name#libaxcd_}
  if [ "$so" = base ]; then unset AXCD_LIB; else export AXCD_LIB=$PWD/$so; fi
  timeout 300 python bench.py --steps 30 --no-next-rows --no-cpu-baseline --no-side-workloads --workload $wl > $out/$name.json 2> $out/$name.err
  python - <<PY
import json
try:
    d=json.load(open("$out/$name.json"))
    print("$name", round(d["ms_per_step"],4), {s["stage"]:s["ms"] for s in d["stages"]})
except Exception as e:
    print("$name", "FAILED", e, open("$out/$name.err").read()[-300:])
PY
done
