#!/bin/bash
# Round 2, multi-GPU visit: real-NCCL slab parity tests (python + C++ host), bench.py under torchrun with c3 / c4.
tag=${1:-r02m}
N=${2:-2}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=index,name --format=csv > $out/smi.txt 2>&1
nvidia-smi topo -m >> $out/smi.txt 2>&1
timeout 900 python -m pytest tests/test_multi_gpu.py tests/test_facade.py -m gpu -q -x -p no:cacheprovider > $out/tests.log 2>&1
echo "pytest exit $?" >> $out/tests.log
tail -15 $out/tests.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 20 --warmup 5 > $out/bench_n$N.json 2> $out/bench_n$N.err
tail -c 1500 $out/bench_n$N.json; tail -5 $out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 3 --warmup 1 > $out/bench_ref_n$N.json 2> $out/bench_ref_n$N.err
ls -la $out
