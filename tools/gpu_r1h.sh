#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TORCH_NCCL_HEARTBEAT_TIMEOUT_SEC=120
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench2.log 2>&1; echo "bench2 rc=$?"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --workload sweep --gpus 2 --steps 10 --warmup 3 > gpurun_out/h_sweep2.log 2>&1; echo "sweep2 rc=$?"
grep -v Warn gpurun_out/h_bench2.log | tail -n 3 | cut -c1-1500;  grep -v Warn gpurun_out/h_sweep2.log | tail -n 2 | cut -c1-400
