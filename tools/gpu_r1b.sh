#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PY="python -m pytest -q -p no:cacheprovider --timeout 600"
timeout 1200 $PY tests/test_gpu_backbone.py -x > gpurun_out/b_tests.log 2>&1; echo "backbone tests rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --profile-kinds --phases --no-cpu-baseline --profile-csv gpurun_out/b_prof.csv > gpurun_out/b_bench.log 2>&1; echo "bench rc=$?"
tail -n 8 gpurun_out/b_tests.log; tail -n 24 gpurun_out/b_bench.log
python tools/prof_summary.py gpurun_out/b_prof.csv 45
