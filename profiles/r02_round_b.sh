#!/bin/bash
# Round 2, visit B (1 GPU): full GPU test suite, the default bench line (graph launch, side workloads), the
# reference arm, the ncu launch list of the bench command.
tag=${1:-r02b}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 1800 python -m pytest tests -m gpu -q -p no:cacheprovider > $out/tests.log 2>&1
echo "pytest exit $?" >> $out/tests.log
tail -15 $out/tests.log
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err
tail -c 300 $out/bench.json; tail -5 $out/bench.err
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $out/bench_ref.json 2> $out/bench_ref.err
AXCD_NO_GRAPH=1 timeout 600 python bench.py --no-next-rows --no-cpu-baseline --no-side-workloads > $out/bench_nograph.json 2> $out/bench_nograph.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-next-rows --no-side-workloads > $out/bench_under_ncu.log 2>&1
ls -la $out
