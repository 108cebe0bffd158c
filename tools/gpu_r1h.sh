#!/bin/bash
# multi-GPU sanity: N-rank training bench (CUDA-graphed head incl. SyncBatchNorm collectives, flat-buffer all-reduce)
N=${1:-2}
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export TORCH_NCCL_HEARTBEAT_TIMEOUT_SEC=120
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/h_bench$N.log 2>&1; echo "bench$N rc=$?"
grep -v Warn gpurun_out/h_bench$N.log | grep '"metric"' | cut -c1-1200
