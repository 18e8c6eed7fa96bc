#!/bin/bash
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l); echo "gpus: $N"; nvidia-smi topo -m | head -14; nproc; lscpu | grep -i "numa\|socket" | head
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus $N --steps 5 --warmup 3 --no-also --no-cpu --e2e-steps 5 > gpurun_out/bench_${N}gpu_e2e.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_${N}gpu_e2e.log | python -c "
import sys,json; d=json.loads(sys.stdin.read()); print('value', round(d['value']), 'e2e', d['e2e'])"
