#!/bin/bash
# Round-2 profile pass (ncu): launch list of the bench command, per-kernel metrics of one training step, metrics of the kernels
# outside the step, one --set full capture of the dominant kernel (stage-1 fc1+GELU NT GEMM).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/p_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/p_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 900 ncu --metrics $M --clock-control none -k regex:"_kernel" -c 2500 --csv --log-file gpurun_out/p_metrics.csv python tools/prof_step.py > gpurun_out/p_ncu2.log 2>&1; echo "step metrics rc=$?"
timeout 600 ncu --metrics $M --clock-control none -k regex:"voxel_|postprocess_|pred2label_|tta_merge_|adamw_|augment_|upload_small|eval_|track_|pack_bbox" -c 400 --csv --log-file gpurun_out/p_small.csv python tools/prof_small.py > gpurun_out/p_ncu4.log 2>&1; echo "small metrics rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gemm_nt_tc_kernel" -s 2 -c 1 -f -o gpurun_out/p_gemm_nt_fc1 python tools/prof_step.py > gpurun_out/p_ncu3.log 2>&1; echo "full capture rc=$?"
ncu -i gpurun_out/p_gemm_nt_fc1.ncu-rep --page details > gpurun_out/p_gemm_nt_fc1_details.txt 2>/dev/null
ncu -i gpurun_out/p_gemm_nt_fc1.ncu-rep --page raw --csv > gpurun_out/p_gemm_nt_fc1_raw.csv 2>/dev/null
tail -n 2 gpurun_out/p_ncu2.log gpurun_out/p_ncu4.log gpurun_out/p_ncu3.log
ls -la gpurun_out/p_*
