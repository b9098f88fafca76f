#!/bin/bash
# Profiling visit (1 GPU): plain bench line, ncu launch list of the same command, --set full captures of one headline
# step and of the C2 narrowphase kernels (direct launches).  Everything lands in gpurun_out/<tag>/.
tag=${1:-r02p}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 900 python bench.py > $out/bench.json 2> $out/bench.err
tail -c 200 $out/bench.json
timeout 600 python bench.py --impl reference --steps 20 --warmup 5 > $out/bench_ref.json 2> $out/bench_ref.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file $out/launches.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-next-rows --no-side-workloads > $out/bench_under_ncu.log 2>&1
AXCD_NO_GRAPH=1 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $out/step \
    python profiles/one_step.py > $out/one_step.log 2>&1
tail -2 $out/one_step.log
AXCD_NO_GRAPH=1 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $out/step_c2 \
    -k regex:"gjkKernel|epaKernel|epaWarpFallbackKernel|classifyPairsKernel|closedFormKernel|slotKernel" python profiles/one_step.py C2 > $out/one_step_c2.log 2>&1
tail -2 $out/one_step_c2.log
timeout 600 python profiles/sort_bandwidth.py > $out/sort_bandwidth.txt 2>&1
ls -la $out
