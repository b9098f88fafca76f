#!/bin/bash
# One GPU-box visit: parity tests, the bench line, the ncu launch list of the same command and a
# `--set full` capture of one step.  Everything lands in gpurun_out/<tag>/.
#   gpurun --timeout 1500 -- 'bash profiles/gpu_round.sh r01x'
tag=${1:-run}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > $out/tests.log 2>&1
echo "pytest exit $?" >> $out/tests.log
tail -5 $out/tests.log
timeout 600 python bench.py > $out/bench.json 2> $out/bench.err
tail -c 600 $out/bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $out/bench_under_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $out/step \
    python profiles/one_step.py > $out/one_step.log 2>&1
tail -2 $out/one_step.log
ls -la $out
