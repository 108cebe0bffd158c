#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PY="python -m pytest -q -p no:cacheprovider --timeout 600"
timeout 1200 $PY tests/test_gpu_kernels.py -k "gemm_tn" -x > gpurun_out/c_tn.log 2>&1; echo "tn tests rc=$?"
timeout 1200 $PY tests/test_gpu_backbone.py > gpurun_out/c_tests.log 2>&1; echo "backbone tests rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --profile-kinds --phases --no-cpu-baseline --profile-csv gpurun_out/c_prof.csv > gpurun_out/c_bench.log 2>&1; echo "bench rc=$?"
tail -n 4 gpurun_out/c_tn.log; tail -n 8 gpurun_out/c_tests.log; tail -n 24 gpurun_out/c_bench.log
python tools/prof_summary.py gpurun_out/c_prof.csv 30
