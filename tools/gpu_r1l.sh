#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PY="python -m pytest -q -p no:cacheprovider --timeout 600"
timeout 1200 $PY tests -m gpu -x --durations=5 > gpurun_out/l_tests.log 2>&1; echo "gpu tests rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/l_smoke.log 2>&1; echo "smoke rc=$?"
timeout 400 python bench.py --steps 10 --warmup 3 --profile-kinds --phases --no-cpu-baseline --profile-csv gpurun_out/l_prof.csv > gpurun_out/l_bench.log 2>&1; echo "bench rc=$?"
tail -n 12 gpurun_out/l_tests.log | cut -c1-200; tail -n 5 gpurun_out/l_smoke.log; grep -v Warning gpurun_out/l_bench.log | tail -n 18 | cut -c1-400
grep "^5," gpurun_out/l_prof.csv | awk -F, '$5>0'
