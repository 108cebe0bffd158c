#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PY="python -m pytest -q -p no:cacheprovider --timeout 300"
timeout 900 $PY tests/test_gpu_kernels.py > gpurun_out/r4_kernels.log 2>&1; echo "kernels rc=$?"
timeout 900 $PY tests/test_gpu_backbone.py -s > gpurun_out/r4_backbone.log 2>&1; echo "backbone rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r4_smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --profile-kinds --no-cpu-baseline --profile-csv gpurun_out/r4_prof.csv > gpurun_out/r4_bench.log 2>&1; echo "bench rc=$?"
tail -n 5 gpurun_out/r4_kernels.log; grep -E "passed|failed|worst|Error|error" gpurun_out/r4_backbone.log | tail -30
tail -n 3 gpurun_out/r4_smoke.log
tail -n 12 gpurun_out/r4_bench.log
python tools/prof_summary.py gpurun_out/r4_prof.csv 40
