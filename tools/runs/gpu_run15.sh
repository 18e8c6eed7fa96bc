#!/bin/bash
# two GPUs: the edge step on NCCL, and both bench arms the way the driver launches them
mkdir -p gpurun_out
nvidia-smi -L
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 tools/shard_check.py > gpurun_out/shard_check.log 2>&1; echo "shard rc=$?"; tail -2 gpurun_out/shard_check.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_2gpu.log 2>&1; echo "bench2 rc=$?"; tail -1 gpurun_out/bench_2gpu.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_2gpu_ref.log 2>&1; echo "bench2 ref rc=$?"; tail -1 gpurun_out/bench_2gpu_ref.log
