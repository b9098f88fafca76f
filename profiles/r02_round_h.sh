#!/bin/bash
# full GPU tests + compute-sanitizer passes over the small all-kernel exercise + headline / C2 bench
tag=${1:-r02h}
out=gpurun_out/$tag
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $out/tests.log 2>&1
echo "pytest exit $?" >> $out/tests.log
tail -6 $out/tests.log
for tool in memcheck racecheck synccheck; do
  AXCD_NO_GRAPH=1 timeout 600 compute-sanitizer --tool $tool --error-exitcode 9 python profiles/sanitize_small.py > $out/san_$tool.log 2>&1
  echo "$tool exit $?" | tee -a $out/san_$tool.log
  grep -E "ERROR SUMMARY|RACECHECK SUMMARY" $out/san_$tool.log | tail -2
done
for wl in C2 headline; do
timeout 300 python bench.py --no-next-rows --no-cpu-baseline --no-side-workloads --workload $wl > $out/bench_$wl.json 2> $out/bench_$wl.err
python - <<PY
import json
d=json.load(open("$out/bench_$wl.json"))
print("$wl", round(d["ms_per_step"],4), {s["stage"]:s["ms"] for s in d["stages"]})
PY
done
