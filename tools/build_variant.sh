#!/bin/bash
# build_variant.sh <tag> <nvcc -D flags...>: a second library with kernels_gemm_tc.cu compiled under extra macros -> leod_b200/lib/libleod_b200_<tag>.so
set -e
HERE="$(cd "$(dirname "$0")/../leod_b200/csrc" && pwd)"
TAG=$1; shift
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $@"
mkdir -p "$HERE/.obj_$TAG"
$NVCC $FLAGS -c "$HERE/kernels_gemm_tc.cu" -o "$HERE/.obj_$TAG/kernels_gemm_tc.o"
OBJS=$(ls "$HERE"/.obj/*.o | grep -v kernels_gemm_tc.o)
$NVCC -shared -o "$HERE/../lib/libleod_b200_$TAG.so" $OBJS "$HERE/.obj_$TAG/kernels_gemm_tc.o" -lcudart
echo "built libleod_b200_$TAG.so"
