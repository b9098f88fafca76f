#!/bin/bash
# multi-GPU bench only (N GPUs): headline replicas + c3 + c4 in one invocation, as the driver runs it
tag=${1:-r02m8}
N=${2:-8}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi topo -m > $out/topo.txt 2>&1
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 \
    bench.py --gpus $N --steps 20 --warmup 5 > $out/bench_n$N.json 2> $out/bench_n$N.err
tail -c 1800 $out/bench_n$N.json; tail -5 $out/bench_n$N.err
ls -la $out
