#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_nt_tc_kernel" -s 3 -c 1 -f -o gpurun_out/k_gemm_gelu python tools/prof_seq.py > gpurun_out/k1.log 2>&1; echo "ncu gelu rc=$?"
ls -la gpurun_out/*.ncu-rep
