#!/bin/bash
# Round 2, visit D (1 GPU): ncu --set full capture of one headline step (direct launches), and of one C2 step.
tag=${1:-r02d}
out=gpurun_out/$tag
mkdir -p $out
AXCD_NO_GRAPH=1 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $out/step \
    python profiles/one_step.py > $out/one_step.log 2>&1
tail -2 $out/one_step.log
AXCD_NO_GRAPH=1 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -f -o $out/step_c2 \
    -k regex:"gjkKernel|epaKernel|epaWarpFallbackKernel|classifyPairsKernel|closedFormKernel|slotKernel" python profiles/one_step.py C2 > $out/one_step_c2.log 2>&1
tail -2 $out/one_step_c2.log
ls -la $out
