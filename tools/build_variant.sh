#!/bin/bash
# Build an experimental variant of the library: tools/build_variant.sh NAME "-DMAGI_KU=5 ..."
# -> gpurun_out/variants/libmagi_NAME.so (use with MAGI_B200_LIB=...; scratch, not part of the product)
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
NAME=$1; shift
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
nvcc $FLAGS $@ -shared -o variants/libmagi_$NAME.so magi_v2_b200/csrc/sampler.cu magi_v2_b200/csrc/cov_build.cu magi_v2_b200/csrc/factor.cu
echo built variants/libmagi_$NAME.so
