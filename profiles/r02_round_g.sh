#!/bin/bash
tag=${1:-r02g}
out=gpurun_out/$tag
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $out/tests.log 2>&1
echo "pytest exit $?" >> $out/tests.log
tail -6 $out/tests.log
for wl in C2 headline; do
timeout 300 python bench.py --no-next-rows --no-cpu-baseline --no-side-workloads --workload $wl > $out/bench_$wl.json 2> $out/bench_$wl.err
python - <<PY
import json
d=json.load(open("$out/bench_$wl.json"))
print("$wl", round(d["ms_per_step"],4), {s["stage"]:s["ms"] for s in d["stages"]})
PY
done
