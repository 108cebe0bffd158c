#!/bin/bash
# quick check: backbone / detect / bench-shape GPU tests + the headline bench line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PY="python -m pytest -q -p no:cacheprovider --timeout 900"
timeout 900 $PY tests/test_gpu_backbone.py tests/test_gpu_bench_shape.py tests/test_gpu_detect.py tests/test_gpu_pseudo_labeler.py -x > gpurun_out/k_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/k_tests.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --profile-kinds --profile-csv gpurun_out/k_prof.csv > gpurun_out/k_bench.log 2>&1; echo "bench rc=$?"
grep -v Warn gpurun_out/k_bench.log | grep "gemm_nt \|metric" | cut -c1-700
