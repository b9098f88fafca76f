#!/bin/bash
# Round 2, visit A: parity tests with the box-box SAT, bench lines (default and box-box through GJK/EPA), C2, launch list.
tag=${1:-r02a}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q -x -p no:cacheprovider > $out/tests.log 2>&1
echo "pytest exit $?" >> $out/tests.log
tail -5 $out/tests.log
timeout 600 python bench.py --no-next-rows > $out/bench.json 2> $out/bench.err
tail -c 300 $out/bench.json
timeout 600 python bench.py --no-next-rows --no-cpu-baseline --flags 8 > $out/bench_generic.json 2> $out/bench_generic.err
timeout 600 python bench.py --no-next-rows --no-cpu-baseline --workload C2 > $out/bench_C2.json 2> $out/bench_C2.err
timeout 600 python bench.py --no-next-rows --no-cpu-baseline --workload C3 > $out/bench_C3.json 2> $out/bench_C3.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-next-rows > $out/bench_under_ncu.log 2>&1
ls -la $out
