#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PY="python -m pytest -q -p no:cacheprovider --timeout 900"
timeout 900 $PY tests/test_gpu_kernels.py tests/test_gpu_backbone.py tests/test_gpu_pseudo_labeler.py tests/test_gpu_detect.py -x > gpurun_out/f_tests.log 2>&1; echo "tests rc=$?"; tail -n 3 gpurun_out/f_tests.log | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-reference-gpu --profile-kinds --profile-csv gpurun_out/f_prof.csv > gpurun_out/f_bench.log 2>&1; echo "bench rc=$?"
grep -v Warn gpurun_out/f_bench.log | grep "launches\|metric" | cut -c1-700
timeout 300 python bench.py --workload sweep --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/f_sweep.log 2>&1; echo "sweep rc=$?"; grep '"metric"' gpurun_out/f_sweep.log | cut -c1-700
