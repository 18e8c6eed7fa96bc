#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "long_delay" > gpurun_out/pytest_long.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_long.log
timeout 300 python tools/sweep.py --workload ns --graph comb --iters 10 --points "mode=exact,fast;layout=planar,interleaved" > gpurun_out/sweep_comb.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --graph comb --iters 10 --points "mode=exact;boxes=1;stages=2,3" >> gpurun_out/sweep_comb.jsonl 2>&1
cat gpurun_out/sweep_comb.jsonl
