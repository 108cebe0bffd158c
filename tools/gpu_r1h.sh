#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TORCH_NCCL_HEARTBEAT_TIMEOUT_SEC=120
timeout 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench2.log 2>&1; echo "bench2 rc=$?"
grep -v Warn gpurun_out/h_bench2.log | tail -n 3 | cut -c1-1500
timeout 240 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench1.log 2>&1; echo "bench1 rc=$?"
grep -v Warn gpurun_out/h_bench1.log | tail -n 1 | cut -c1-900
