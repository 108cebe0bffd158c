#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PY="python -m pytest -q -p no:cacheprovider --timeout 600"
timeout 600 $PY tests/test_gpu_kernels.py -x -k "gemm" > gpurun_out/n_tests.log 2>&1; echo "gemm tests rc=$?"; tail -n 2 gpurun_out/n_tests.log | cut -c1-200
timeout 120 python tools/nt_probe.py 2>&1 | grep "M=" | tee gpurun_out/n_probe.log
