#!/bin/bash
# epilogue-warp variants of the NT GEMM: bench each library build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for tag in "" _w12 _g16 _w12g16; do
  LEOD_B200_LIB=$PWD/leod_b200/lib/libleod_b200$tag.so timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --no-reference-gpu --profile-kinds > gpurun_out/e_var$tag.log 2>&1
  echo "variant '$tag' rc=$?"; grep -v Warn gpurun_out/e_var$tag.log | grep "gemm_nt \|metric" | cut -c1-330
done
PY="python -m pytest -q -p no:cacheprovider --timeout 900"
timeout 600 $PY tests/test_gpu_kernels.py tests/test_gpu_pseudo_labeler.py -x > gpurun_out/e_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/e_tests.log
