#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
PY="python -m pytest -q -p no:cacheprovider --timeout 300"
timeout 600 $PY tests/test_gpu_kernels.py -k "layernorm or lstm_gates" > gpurun_out/r2_ln_lstm.log 2>&1; echo "ln/lstm rc=$?"
timeout 600 $PY tests/test_gpu_kernels.py -k "gemm_tn" > gpurun_out/r2_tn.log 2>&1; echo "tn rc=$?"
timeout 600 $PY tests/test_gpu_kernels.py -k "attention" > gpurun_out/r2_attn.log 2>&1; echo "attn rc=$?"
timeout 600 python tests/gpu_diag.py > gpurun_out/r2_diag.log 2>&1; echo "diag rc=$?"
timeout 600 python bench.py --steps 5 --warmup 3 --profile-kinds --no-cpu-baseline > gpurun_out/r2_bench.log 2>&1; echo "bench rc=$?"
tail -n 25 gpurun_out/r2_ln_lstm.log; tail -n 25 gpurun_out/r2_tn.log; tail -n 25 gpurun_out/r2_attn.log
cat gpurun_out/r2_diag.log | tail -100
tail -n 12 gpurun_out/r2_bench.log
