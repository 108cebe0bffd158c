#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PY="python -m pytest -q -p no:cacheprovider --timeout 300"
timeout 600 $PY tests/test_gpu_kernels.py -k "attention" -x > gpurun_out/i_attn.log 2>&1; echo "attention tests rc=$?"
timeout 600 $PY tests/test_gpu_backbone.py -x > gpurun_out/i_tests.log 2>&1; echo "backbone tests rc=$?"
timeout 400 python bench.py --steps 10 --warmup 3 --profile-kinds --phases --no-cpu-baseline --profile-csv gpurun_out/i_prof.csv > gpurun_out/i_bench.log 2>&1; echo "bench rc=$?"
tail -n 4 gpurun_out/i_attn.log; tail -n 4 gpurun_out/i_tests.log; grep -v Warning gpurun_out/i_bench.log | tail -n 18 | cut -c1-400
python tools/prof_summary.py gpurun_out/i_prof.csv 14
