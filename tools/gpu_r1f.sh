#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PY="python -m pytest -q -p no:cacheprovider --timeout 600"
timeout 1200 $PY tests/test_gpu_kernels.py -x > gpurun_out/f_kernels.log 2>&1; echo "kernel tests rc=$?"
timeout 1200 $PY tests/test_gpu_backbone.py -x > gpurun_out/f_tests.log 2>&1; echo "backbone tests rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --profile-kinds --phases --no-cpu-baseline --profile-csv gpurun_out/f_prof.csv > gpurun_out/f_bench.log 2>&1; echo "bench rc=$?"
tail -n 4 gpurun_out/f_kernels.log; tail -n 4 gpurun_out/f_tests.log; grep -v Warning gpurun_out/f_bench.log | tail -n 18 | cut -c1-400
python tools/prof_summary.py gpurun_out/f_prof.csv 32
