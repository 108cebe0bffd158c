#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PY="python -m pytest -q -p no:cacheprovider --timeout 600"
timeout 600 $PY tests/test_gpu_augment.py tests/test_gpu_eval_steps.py tests/test_gpu_self_training.py -x > gpurun_out/b_tests.log 2>&1; echo "tests rc=$?"
timeout 600 python bench.py --no-cpu-baseline --no-reference-gpu > gpurun_out/b_bench.log 2>&1; echo "bench rc=$?"
tail -n 25 gpurun_out/b_tests.log | cut -c1-300; grep '"metric"' gpurun_out/b_bench.log | cut -c1-900
