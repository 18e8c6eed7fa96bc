#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 tools/shard_check.py > gpurun_out/shard_check_p2p_$N.log 2>&1; echo "shard rc=$?"; tail -12 gpurun_out/shard_check_p2p_$N.log | cut -c1-600
