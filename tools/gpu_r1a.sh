#!/bin/bash
# Re-entry GPU pass: full GPU suite, smoke, bench (with per-launch CSV), reference arm, ncu launch list.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
nproc > gpurun_out/nproc.txt
PY="python -m pytest -q -p no:cacheprovider --timeout 600"
timeout 1200 $PY tests -m gpu -x > gpurun_out/a_tests.log 2>&1; echo "gpu tests rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/a_smoke.log 2>&1; echo "smoke rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 --profile-kinds --profile-csv gpurun_out/a_prof.csv > gpurun_out/a_bench.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/a_ref.log 2>&1; echo "ref rc=$?"
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/a_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/a_ncu_bench.log 2>&1; echo "ncu rc=$?"
tail -n 5 gpurun_out/a_tests.log; tail -n 3 gpurun_out/a_smoke.log; tail -n 14 gpurun_out/a_bench.log; tail -n 2 gpurun_out/a_ref.log
python tools/prof_summary.py gpurun_out/a_prof.csv 40
