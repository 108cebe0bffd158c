#!/bin/bash
# Round-end validation: full GPU suite, smoke, every bench workload, both reference arms, per-launch CSV of the headline step.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PY="python -m pytest -q -p no:cacheprovider --timeout 900"
timeout 1500 $PY tests -m gpu > gpurun_out/z_tests.log 2>&1; echo "gpu tests rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/z_smoke.log 2>&1; echo "smoke rc=$?"
timeout 400 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/z_ref.log 2>&1; echo "ref rc=$?"
timeout 900 python bench.py --profile-kinds --phases --profile-csv gpurun_out/z_prof.csv > gpurun_out/z_bench.log 2>&1; echo "bench rc=$?"
for w in train-dense train-gen4 selftrain; do
  timeout 600 python bench.py --workload $w --steps 8 --warmup 3 --profile-kinds --no-cpu-baseline > gpurun_out/z_bench_$w.log 2>&1; echo "$w rc=$?"
done
timeout 400 python bench.py --workload sweep --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/z_sweep.log 2>&1; echo "sweep rc=$?"
timeout 120 python tools/nt_probe.py > gpurun_out/z_nt_probe.log 2>&1
tail -n 3 gpurun_out/z_tests.log | cut -c1-200; grep "smoke" gpurun_out/z_smoke.log | cut -c1-300; grep '"impl"' gpurun_out/z_ref.log | cut -c1-300
grep -v Warn gpurun_out/z_bench.log | grep "launches\|phase\|metric" | cut -c1-1500
for w in train-dense train-gen4 selftrain; do grep '"metric"' gpurun_out/z_bench_$w.log | cut -c1-420; done
grep '"metric"' gpurun_out/z_sweep.log | cut -c1-700
