#!/bin/bash
# Runs bench.py once per library variant under axiom-physics-engine_b200/variants/ and prints the stage table.
# usage (on the GPU box): bash profiles/variant_sweep.sh A B C ...
mkdir -p gpurun_out
for t in "$@"; do
  AXCD_LIB=$PWD/axiom-physics-engine_b200/variants/libaxcd_$t.so timeout 300 python bench.py --steps 60 --warmup 5 --no-cpu-baseline --no-next-rows > gpurun_out/sweep_$t.json 2> gpurun_out/sweep_$t.err
  python - "$t" <<'PY'
import json,sys
t=sys.argv[1]
try:
    d=json.loads(open(f"gpurun_out/sweep_{t}.json").read().strip().splitlines()[-1])
    print(t, "ms/step", round(d["ms_per_step"],4), d['config']['candidate_pairs'], d['config']['contacts'], [(s["stage"], s["ms"]) for s in d["stages"]])
except Exception as e:
    print(t, "FAILED", e, open(f"gpurun_out/sweep_{t}.err").read()[-400:])
PY
done
