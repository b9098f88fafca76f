#!/bin/bash
# bottom-up LBVH build A/B: full GPU tests on the default (bottom-up) path, the parity file again on the top-down
# path, then headline / C2 bench lines for both
tag=${1:-r02p2}
out=gpurun_out/$tag
mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -p no:cacheprovider > $out/tests.log 2>&1
echo "pytest exit $?" >> $out/tests.log
tail -6 $out/tests.log
AXCD_BUILD_TOPDOWN=1 timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -p no:cacheprovider > $out/tests_topdown.log 2>&1
echo "pytest(topdown) exit $?" >> $out/tests_topdown.log
tail -3 $out/tests_topdown.log
for mode in bottomup topdown; do
  if [ $mode = topdown ]; then export AXCD_BUILD_TOPDOWN=1; else unset AXCD_BUILD_TOPDOWN; fi
  for wl in headline C2; do
    timeout 300 python bench.py --no-next-rows --no-cpu-baseline --no-side-workloads --workload $wl > $out/bench_${wl}_$mode.json 2> $out/bench_${wl}_$mode.err
    python - <<PY
import json
try:
    d=json.load(open("$out/bench_${wl}_$mode.json"))
    print("$wl $mode", round(d["ms_per_step"],4), {s["stage"]:s["ms"] for s in d["stages"]})
except Exception as e:
    print("$wl $mode FAILED", e, open("$out/bench_${wl}_$mode.err").read()[-400:])
PY
  done
done
