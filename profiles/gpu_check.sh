#!/bin/bash
# Quick GPU visit: the -m gpu suite, then one bench line (no CPU baseline) with the stage table printed.
#   gpurun --timeout 1200 -- 'bash profiles/gpu_check.sh <tag> [extra bench args]'
tag=${1:-check}; shift
out=gpurun_out/$tag; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $out/tests.log 2>&1; echo "pytest exit $?" >> $out/tests.log; tail -12 $out/tests.log
timeout 600 python bench.py --no-cpu-baseline "$@" > $out/bench.json 2> $out/bench.err; tail -3 $out/bench.err
python - $out <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]+'/bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['config']['candidate_pairs'], d['config']['contacts'], [(s['stage'], s['ms'], s['frac_of_hbm_peak']) for s in d['stages']])
print('e2e', d['e2e']['ms_per_step'], d.get('next_rows'))
PY
