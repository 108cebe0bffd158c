#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,lts__t_sectors_op_write.sum,lts__t_sectors_op_read.sum,smsp__inst_executed.sum,sm__inst_executed_pipe_xu.sum,smsp__cycles_active.avg,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"
for skip in 7 13; do
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_nt_tc_kernel -s $skip -c 1 -f -o gpurun_out/e_gemm_nt_$skip python tools/prof_seq.py > gpurun_out/e_ncu_$skip.log 2>&1; echo "ncu gemm $skip rc=$?"
done
timeout 900 ncu --metrics $M --clock-control none -k regex:"gemm_nt_tc|ln_fwd|ln_bwd|attn_bwd|attn_fwd|im2col|col2im|lstm_fwd|lstm_bwd|gemm_tn_tc|colsum" -c 600 --csv --log-file gpurun_out/e_metrics.csv python tools/prof_seq.py bwd > gpurun_out/e_ncu2.log 2>&1; echo "ncu metrics rc=$?"
ls -la gpurun_out/
