#!/bin/bash
# Round-end validation: full GPU suite, smoke, bench (both arms), per-launch CSV, short ncu launch list of one bench step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PY="python -m pytest -q -p no:cacheprovider --timeout 600"
timeout 1200 $PY tests -m gpu > gpurun_out/z_tests.log 2>&1; echo "gpu tests rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1; echo "smoke rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/z_ref.log 2>&1; echo "ref rc=$?"
timeout 600 python bench.py --profile-kinds --phases --profile-csv gpurun_out/z_prof.csv > gpurun_out/z_bench.log 2>&1; echo "bench rc=$?"
timeout 300 python bench.py --workload sweep --steps 10 --warmup 3 > gpurun_out/z_sweep.log 2>&1; echo "sweep rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^((?!batchnorm).)*$' -s 7600 -c 1500 --csv --log-file gpurun_out/z_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/z_ncu_bench.log 2>&1; echo "launch list rc=$?"
tail -n 3 gpurun_out/z_tests.log | cut -c1-200; tail -n 2 gpurun_out/z_smoke.log; grep '"impl"' gpurun_out/z_ref.log | cut -c1-300; grep -v Warn gpurun_out/z_bench.log | tail -n 18 | cut -c1-1800; grep '"metric"' gpurun_out/z_sweep.log | cut -c1-300
