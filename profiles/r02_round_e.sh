#!/bin/bash
# Round 2, visit E: traversal with dense leaf tests — parity tests + A/B bench
tag=${1:-r02e}
out=gpurun_out/$tag
mkdir -p $out
timeout 1800 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $out/tests.log 2>&1
echo "pytest exit $?" >> $out/tests.log
tail -8 $out/tests.log
timeout 600 python bench.py --no-next-rows --no-cpu-baseline --no-side-workloads > $out/bench.json 2> $out/bench.err
AXCD_TRAV_INLINE=1 timeout 600 python bench.py --no-next-rows --no-cpu-baseline --no-side-workloads > $out/bench_inline.json 2> $out/bench_inline.err
timeout 600 python bench.py --no-next-rows --no-cpu-baseline --no-side-workloads --workload C3 > $out/bench_C3.json 2> $out/bench_C3.err
timeout 600 python bench.py --no-next-rows --no-cpu-baseline --no-side-workloads --workload C2 > $out/bench_C2.json 2> $out/bench_C2.err
for f in bench bench_inline bench_C3 bench_C2; do python - <<PY
import json
d=json.load(open("$out/$f.json"))
print("$f", round(d["ms_per_step"],4), {s["stage"]:s["ms"] for s in d["stages"]})
PY
done
