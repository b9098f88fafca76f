#!/bin/bash
# multi-GPU parity tests only (short timeouts: a hang must not burn the budget)
tag=${1:-r02mt}
out=gpurun_out/$tag
mkdir -p $out
timeout 300 python -m pytest tests/test_multi_gpu.py -m gpu -q -x -p no:cacheprovider > $out/tests_py.log 2>&1
echo "pytest exit $?" >> $out/tests_py.log
tail -15 $out/tests_py.log
timeout 240 python -m pytest tests/test_facade.py -m gpu -q -x -p no:cacheprovider -k slab_host > $out/tests_cpp.log 2>&1
echo "pytest exit $?" >> $out/tests_cpp.log
tail -15 $out/tests_cpp.log
timeout 120 axiom-physics-engine_b200/slab_host 400000 73.7 7 2 6 > $out/slab_host.txt 2>&1; echo "slab_host exit $?" >> $out/slab_host.txt; cat $out/slab_host.txt
timeout 300 python bench.py --no-next-rows --no-cpu-baseline --no-side-workloads --workload C2 > $out/bench_C2.json 2> $out/bench_C2.err
python - <<PY
import json
d=json.load(open("$out/bench_C2.json"))
print("C2", round(d["ms_per_step"],4), {s["stage"]:s["ms"] for s in d["stages"]})
PY
