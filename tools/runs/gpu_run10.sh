#!/bin/bash
# K1b with two chunks per iteration; refill policy
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
timeout 300 python tools/sweep.py --workload c2 --points "mode=exact,fast;lanes=4" > gpurun_out/sweep_k1b.jsonl 2>&1
timeout 300 python tools/sweep.py --workload c8k --points "mode=exact;lanes=4,1" >> gpurun_out/sweep_k1b.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --points "mode=exact,fast" >> gpurun_out/sweep_k1b.jsonl 2>&1
cat gpurun_out/sweep_k1b.jsonl
