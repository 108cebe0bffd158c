#!/bin/bash
# final-state refresh of the ncu summaries: launch list of the bench command, per-kernel metrics of one step, small kernels
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
M="gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/q_launches.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-reference-gpu > gpurun_out/q_ncu_bench.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --metrics $M --clock-control none -k regex:"_kernel" -c 2500 --csv --log-file gpurun_out/q_metrics.csv python tools/prof_step.py > gpurun_out/q_ncu2.log 2>&1; echo "step metrics rc=$?"
timeout 400 ncu --metrics $M --clock-control none -k regex:"voxel_|postprocess_|pred2label_|tta_merge_|adamw_|augment_|upload_small|eval_|track_|pack_bbox" -c 400 --csv --log-file gpurun_out/q_small.csv python tools/prof_small.py > gpurun_out/q_ncu4.log 2>&1; echo "small metrics rc=$?"
ls -la gpurun_out/q_*
