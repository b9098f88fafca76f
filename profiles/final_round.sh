#!/bin/bash
# Final artefacts of a round: bench lines of the named workloads (not under a profiler).
out=gpurun_out/${1:-final}; mkdir -p $out
for wl in C1 C2 C3; do
  timeout 600 python bench.py --workload $wl --steps 30 > $out/bench_$wl.json 2> $out/bench_$wl.err
  python - $out/bench_$wl.json <<'PY'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1]); print(d['config']['workload'][:40], d['ms_per_step'], d['value'], d.get('cpu_baseline',{}).get('ms_per_step'), d.get('cpu_baseline',{}).get('pairs_match_gpu'))
PY
done
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $out/bench_reference_arm.json 2>&1; tail -c 400 $out/bench_reference_arm.json
python profiles/sort_bandwidth.py > $out/sort_bandwidth.txt 2>&1; tail -8 $out/sort_bandwidth.txt
