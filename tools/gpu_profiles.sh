#!/bin/bash
# Round profile pass: ncu launch list of the bench command, per-kernel DRAM/tensor metrics of one backbone fwd+bwd,
# one --set full capture of the dominant kernel class.  Outputs stay small (CSV/text), see profiles/.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"
# cuDNN's semi-persistent BatchNorm kernels are cooperative launches that ncu cannot serialise: leave them out of the list
timeout 1200 ncu --metrics gpu__time_duration.sum --clock-control none -k 'regex:^((?!batchnorm).)*$' -c 14000 --csv --log-file gpurun_out/p_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > gpurun_out/p_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --metrics $M --clock-control none -k regex:"gemm_nt_tc|gemm_tn_tc|colsum|ln_fwd|ln_bwd|attn_bwd|attn_fwd|im2col|col2im|lstm_" -c 700 --csv --log-file gpurun_out/p_metrics.csv python tools/prof_seq.py bwd > gpurun_out/p_ncu2.log 2>&1; echo "metrics rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_nt_tc_kernel" -s 3 -c 1 -f -o gpurun_out/p_gemm_nt_fc1 python tools/prof_seq.py > gpurun_out/p_ncu3.log 2>&1; echo "full capture rc=$?"
ncu -i gpurun_out/p_gemm_nt_fc1.ncu-rep --page details > gpurun_out/p_gemm_nt_fc1_details.txt 2>/dev/null
ls -la gpurun_out/p_*
