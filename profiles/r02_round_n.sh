#!/bin/bash
# parity (full GPU suite), A/B of the headline against the previous library build, and the full bench line
tag=${1:-r02n}
out=gpurun_out/$tag
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $out/tests.log 2>&1
echo "pytest exit $?" >> $out/tests.log
tail -5 $out/tests.log
bash profiles/variant_bench.sh $tag/var headline
timeout 900 python bench.py --no-cpu-baseline --no-next-rows --no-side-workloads > $out/bench.json 2> $out/bench.err
python - <<PY
import json
d=json.load(open("$out/bench.json"))
print("step", round(d["ms_per_step"],4), "e2e", d["e2e"], "full", d["e2e_full_transforms"])
PY
