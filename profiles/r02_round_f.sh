#!/bin/bash
# 1 GPU: full GPU suite (cylinder, fused narrow, ...), default bench
tag=${1:-r02f}
out=gpurun_out/$tag
mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider > $out/tests.log 2>&1
echo "pytest exit $?" >> $out/tests.log
tail -12 $out/tests.log
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err
python - <<PY
import json
d=json.load(open("$out/bench.json"))
print(round(d["ms_per_step"],4), "e2e", round(d["e2e"]["ms_per_step"],3), {s["stage"]:s["ms"] for s in d["stages"]})
print("c2", d["c2"]["ms_per_step"], "generic", d["boxbox_through_gjk_epa"]["ms_per_step"])
PY
