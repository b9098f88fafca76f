#!/bin/bash
# Multi-GPU bench lines on one box: gpurun --gpus N -- 'bash profiles/multi_gpu.sh N <tag>'
N=${1:-2}; out=gpurun_out/${2:-mgpu}; mkdir -p $out
run() { name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N "$@" > $out/$name.json 2> $out/$name.err
  python - $out/$name.json <<'PY'
import json,sys
try:
    d=json.loads([l for l in open(sys.argv[1]).read().strip().splitlines() if l.startswith('{')][-1])
    print(d['config'].get('workload','')[:50], 'n_gpus', d['n_gpus'], 'ms/step', round(d['ms_per_step'],4), 'value', d['value'], 'e2e', (d.get('e2e') or {}).get('value'))
except Exception as e:
    print('FAILED', sys.argv[1], e)
PY
}
run bench_n${N} --steps 30 --warmup 3
run bench_n${N}_C3 --workload C3 --steps 30 --warmup 3
run bench_n${N}_C4 --workload C4 --scale 0.25 --steps 10 --warmup 3
