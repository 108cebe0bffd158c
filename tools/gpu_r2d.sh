#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PY="python -m pytest -q -p no:cacheprovider --timeout 900"
timeout 1200 $PY tests/test_gpu_kernels.py tests/test_gpu_backbone.py tests/test_gpu_detect.py tests/test_gpu_bench_shape.py tests/test_gpu_augment.py -x > gpurun_out/d_tests.log 2>&1; echo "tests rc=$?"
timeout 600 python bench.py --no-cpu-baseline --no-reference-gpu --profile-kinds > gpurun_out/d_bench.log 2>&1; echo "bench rc=$?"
tail -n 6 gpurun_out/d_tests.log | cut -c1-300; grep -v Warn gpurun_out/d_bench.log | grep "launches\|metric" | cut -c1-900
