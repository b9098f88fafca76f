#!/bin/bash
# e2e A/B on one box: contact sink filled by stores from the SMs (variants/libaxcd_zc.so, built from the previous commit)
# against the copy-engine chunks that follow the narrowphase's progress words, over chunk-size settings.
out=gpurun_out/${1:-r03c}; mkdir -p $out
run() { name=$1; shift; env "$@" timeout 200 python bench.py --steps 40 --no-next-rows --no-cpu-baseline --no-side-workloads > $out/$name.json 2> $out/$name.err; python -c "
import json; d=json.load(open('$out/$name.json')); print('$name', d['ms_per_step'], d['e2e']['ms_per_step'], d['e2e_full_transforms']['ms_per_step'])"; }
[ -e axiom-physics-engine_b200/variants/libaxcd_zc.so ] && run zc AXCD_LIB=$PWD/axiom-physics-engine_b200/variants/libaxcd_zc.so
run dflt X=1
run fixed8 AXCD_SINK_CHUNK_MIN_KB=8192
run min64 AXCD_SINK_CHUNK_MIN_KB=64
run min1024 AXCD_SINK_CHUNK_MIN_KB=1024
run max16 AXCD_SINK_CHUNK_KB=16384
run dflt2 X=1
