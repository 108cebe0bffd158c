#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PY="python -m pytest -q -p no:cacheprovider --timeout 300"
timeout 600 $PY tests/test_gpu_pseudo_labeler.py -x > gpurun_out/g_pl.log 2>&1; echo "pseudo-labeler tests rc=$?"
timeout 300 python bench.py --workload sweep --steps 10 --warmup 3 > gpurun_out/g_sweep.log 2>&1; echo "sweep rc=$?"
tail -n 6 gpurun_out/g_pl.log; grep -v Warn gpurun_out/g_sweep.log | tail -n 3 | cut -c1-900
