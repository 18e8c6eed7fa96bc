#!/bin/bash
# eight GPUs: edge step on NCCL and both bench arms the way the driver launches them
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/shard_check.py > gpurun_out/shard_check_$N.log 2>&1; echo "shard rc=$?"; tail -1 gpurun_out/shard_check_$N.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 20 --warmup 3 > gpurun_out/bench_${N}gpu.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_${N}gpu.log | cut -c1-1500
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519 bench.py --impl reference --gpus $N --steps 3 --warmup 1 > gpurun_out/bench_${N}gpu_ref.log 2>&1; echo "bench ref rc=$?"; tail -1 gpurun_out/bench_${N}gpu_ref.log | cut -c1-400
