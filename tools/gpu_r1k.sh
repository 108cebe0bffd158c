#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_nt_tc_kernelILi1" -c 1 -f -o gpurun_out/k_gemm_gelu python tools/prof_seq.py > gpurun_out/k1.log 2>&1; echo "ncu gelu rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"lstm_seq_fwd" -c 1 -f -o gpurun_out/k_lstm python tools/prof_seq.py > gpurun_out/k2.log 2>&1; echo "ncu lstm rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"attn_bwd" -s 6 -c 1 -f -o gpurun_out/k_attn_bwd python tools/prof_seq.py bwd > gpurun_out/k3.log 2>&1; echo "ncu attn rc=$?"
ls -la gpurun_out/*.ncu-rep
