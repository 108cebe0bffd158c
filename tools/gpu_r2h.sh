#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/h_smoke.log 2>&1; echo "smoke rc=$?"; grep "smoke" gpurun_out/h_smoke.log | cut -c1-250
