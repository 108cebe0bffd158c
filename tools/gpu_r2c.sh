#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PY="python -m pytest -q -p no:cacheprovider --timeout 900"
timeout 900 $PY tests/test_gpu_eval_metrics.py -x > gpurun_out/c_tests.log 2>&1; echo "tests rc=$?"
timeout 400 python bench.py --workload sweep --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/c_sweep.log 2>&1; echo "sweep rc=$?"
tail -n 40 gpurun_out/c_tests.log | cut -c1-400; grep '"metric"' gpurun_out/c_sweep.log | cut -c1-1100
