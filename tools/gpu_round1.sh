#!/bin/bash
# First GPU bring-up: kernels in isolation (safe ones first), then the backbone, smoke and a short bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
PY="python -m pytest -q -p no:cacheprovider --timeout 300"
timeout 900 $PY tests/test_gpu_kernels.py -k "not tcgen05 and not two_sources and not gemm_tn" > gpurun_out/t1_kernels_safe.log 2>&1; echo "safe kernels rc=$?"
timeout 600 $PY tests/test_gpu_kernels.py -k "gemm_tn or two_sources" > gpurun_out/t2_tn.log 2>&1; echo "tn/two_sources rc=$?"
timeout 600 $PY tests/test_gpu_kernels.py -k "tcgen05" > gpurun_out/t3_tcgen05.log 2>&1; echo "tcgen05 rc=$?"
timeout 900 $PY tests/test_gpu_backbone.py -k "fp32 or accumulation or reset" > gpurun_out/t4_backbone_fp32.log 2>&1; echo "backbone fp32 rc=$?"
timeout 900 $PY tests/test_gpu_backbone.py -k "bf16" > gpurun_out/t5_backbone_bf16.log 2>&1; echo "backbone bf16 rc=$?"
timeout 900 $PY tests/test_gpu_backbone.py -k "full_size" > gpurun_out/t6_fullsize.log 2>&1; echo "fullsize rc=$?"
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"
timeout 600 python bench.py --steps 3 --warmup 3 --gemm-impl 0 --profile-kinds --no-cpu-baseline > gpurun_out/bench_simt.log 2>&1; echo "bench simt rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --profile-kinds > gpurun_out/bench_tc.log 2>&1; echo "bench tc rc=$?"
tail -n 3 gpurun_out/t*.log gpurun_out/smoke.log
tail -n 12 gpurun_out/bench_simt.log gpurun_out/bench_tc.log
