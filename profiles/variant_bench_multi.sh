#!/bin/bash
# Benchmarks the default library and every variants/libaxcd_*.so on several workloads: one line per (variant, workload).
tag=${1:-var}
shift
out=gpurun_out/$tag
mkdir -p $out
for so in base axiom-physics-engine_b200/variants/libaxcd_*.so; do
  [ -e "$so" ] || [ "$so" = base ] || continue
  name=$(basename $so .so); name=${name#libaxcd_}
  if [ "$so" = base ]; then unset AXCD_LIB; else export AXCD_LIB=$PWD/$so; fi
  for wl in "$@"; do
    timeout 300 python bench.py --steps 30 --no-next-rows --no-cpu-baseline --no-side-workloads --workload $wl > $out/${name}_$wl.json 2> $out/${name}_$wl.err
    python - <<PY
import json
try:
    d=json.load(open("$out/${name}_$wl.json"))
    print("$name $wl", round(d["ms_per_step"],4), {s["stage"]:s["ms"] for s in d["stages"]})
except Exception as e:
    print("$name $wl", "FAILED", e, open("$out/${name}_$wl.err").read()[-300:])
PY
  done
done
