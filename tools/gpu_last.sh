#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PY="python -m pytest -q -p no:cacheprovider --timeout 900"
timeout 1200 $PY tests -m gpu > gpurun_out/y_tests.log 2>&1; echo "gpu tests rc=$?"; tail -n 2 gpurun_out/y_tests.log | cut -c1-200
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/y_smoke.log 2>&1; echo "smoke rc=$?"; grep "smoke" gpurun_out/y_smoke.log | tail -2 | cut -c1-200
timeout 600 python bench.py --profile-kinds --phases --profile-csv gpurun_out/y_prof.csv > gpurun_out/y_bench.log 2>&1; echo "bench rc=$?"
grep -v Warn gpurun_out/y_bench.log | grep "launches\|phase\|metric" | cut -c1-1200
