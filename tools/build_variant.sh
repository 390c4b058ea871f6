#!/bin/bash
# Build an experimental variant of the library: tools/build_variant.sh NAME "-DMAGI_TS_C=3 ..." [SRCROOT]
# -> variants/libmagi_NAME.so (use with MAGI_B200_LIB=...; scratch, git-ignored, not part of the product).
# Only one translation unit -- sampler.cu (the fast-path kernels) unless VARIANT_TU names another, e.g.
# VARIANT_TU=factor -- is recompiled with the extra flags; the other objects are the in-tree ones
# (magi_v2_b200/build/*.o, made by `python -m magi_v2_b200.build`).  -DMAGI_DEV_SEIR4_ONLY compiles one model only.
set -e
cd "$(dirname "$0")/.."
mkdir -p variants
NAME=$1; FLAGS_X=$2; SRC=${3:-.}
FLAGS="-gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden"
TU=${VARIANT_TU:-sampler}
nvcc $FLAGS $FLAGS_X -c -o variants/${TU}_$NAME.o $SRC/magi_v2_b200/csrc/$TU.cu
OTHERS=$(ls magi_v2_b200/build/*.o | grep -v "/$TU.o")
nvcc -gencode arch=compute_100a,code=sm_100a -shared -o variants/libmagi_$NAME.so variants/${TU}_$NAME.o $OTHERS
rm -f variants/${TU}_$NAME.o
echo built variants/libmagi_$NAME.so
