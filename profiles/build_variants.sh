#!/bin/bash
# Builds tuning variants of libaxcd.so into axiom-physics-engine_b200/variants/ (git-ignored; they travel to the GPU box).
#   profiles/build_variants.sh NAME "-DFLAG=.. -DFLAG2=.." [NAME2 "..."]
cd "$(dirname "$0")/../axiom-physics-engine_b200" || exit 1
mkdir -p variants
while [ $# -ge 2 ]; do
  name=$1; flags=$2; shift 2
  /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -std=c++20 -O3 -lineinfo -fmad=false -prec-div=true -prec-sqrt=true -ftz=false \
     -Xcompiler -fPIC,-fvisibility=hidden -cudart static -I../include $flags -shared -o variants/libaxcd_$name.so csrc/axcd_api.cu &
done
wait
ls -la variants/
