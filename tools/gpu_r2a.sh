#!/bin/bash
# mid-round check: GPU suite + bench with per-class profile
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PY="python -m pytest -q -p no:cacheprovider --timeout 600"
timeout 1200 $PY tests -m gpu -x > gpurun_out/a_tests.log 2>&1; echo "gpu tests rc=$?"
timeout 600 python bench.py --profile-kinds --phases --profile-csv gpurun_out/a_prof.csv --no-cpu-baseline > gpurun_out/a_bench.log 2>&1; echo "bench rc=$?"
tail -n 3 gpurun_out/a_tests.log | cut -c1-300; grep -v Warn gpurun_out/a_bench.log | tail -n 30 | cut -c1-2500
