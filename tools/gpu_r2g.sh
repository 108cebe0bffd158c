#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for d in 0 3; do LEOD_NT_DEBUG=$d timeout 120 python tools/nt_probe.py 2>&1 | grep "dbg="; done | tee gpurun_out/g_probe5.log
PY="python -m pytest -q -p no:cacheprovider --timeout 900"
timeout 600 $PY tests/test_gpu_kernels.py -x -k "gemm" > gpurun_out/g_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/g_tests.log | cut -c1-200
