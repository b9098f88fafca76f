#!/bin/bash
# Round 2, visit C (1 GPU): GPU tests, default bench (fused closed-form narrowphase), A/B with the split kernels,
# sort micro-bench (LSD vs bucket sort), launch list.
tag=${1:-r02c}
out=gpurun_out/$tag
mkdir -p $out
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > $out/tests.log 2>&1
echo "pytest exit $?" >> $out/tests.log
tail -8 $out/tests.log
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err
tail -c 200 $out/bench.json; tail -3 $out/bench.err
AXCD_SPLIT_NARROW=1 timeout 600 python bench.py --no-next-rows --no-cpu-baseline --no-side-workloads > $out/bench_split.json 2> $out/bench_split.err
AXCD_BUCKET_SORT=1 timeout 600 python bench.py --no-next-rows --no-cpu-baseline --no-side-workloads > $out/bench_bucket.json 2> $out/bench_bucket.err
timeout 600 python profiles/sort_bandwidth.py > $out/sort_bandwidth.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-next-rows --no-side-workloads > $out/bench_under_ncu.log 2>&1
ls -la $out
