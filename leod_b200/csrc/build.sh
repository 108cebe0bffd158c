#!/bin/bash
# Builds leod_b200/lib/libleod_b200.so for sm_100a.  Usage: build.sh [extra nvcc flags]
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../lib"
mkdir -p "$OUT" "$HERE/.obj"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
FLAGS="-gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC -Xcompiler -Wall --expt-relaxed-constexpr $@"
pids=()
for f in "$HERE"/*.cu; do
  o="$HERE/.obj/$(basename "${f%.cu}").o"
  if [ ! -f "$o" ] || [ "$f" -nt "$o" ] || [ "$HERE/common.cuh" -nt "$o" ] || [ "$HERE/tc_ptx.cuh" -nt "$o" ] || [ "$HERE/../../include/leod_b200.h" -nt "$o" ]; then
    $NVCC $FLAGS -c "$f" -o "$o" &
    pids+=($!)
  fi
done
for p in "${pids[@]}"; do wait $p; done
$NVCC -shared -o "$OUT/libleod_b200.so" "$HERE"/.obj/*.o -lcudart
echo "built $OUT/libleod_b200.so"
