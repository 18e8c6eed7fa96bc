#!/bin/bash
mkdir -p gpurun_out
./tools/ubench/f32x2 > gpurun_out/ubench_f32x2.txt 2>&1; cat gpurun_out/ubench_f32x2.txt
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest.log
timeout 300 python tools/sweep.py --workload c2 --points "mode=exact,fast;lanes=1,4" > gpurun_out/sweep2.jsonl 2>&1
timeout 300 python tools/sweep.py --workload c2 --points "mode=exact,fast;lanes=4;boxes=4,16;stages=2,4" >> gpurun_out/sweep2.jsonl 2>&1
timeout 300 python tools/sweep.py --workload c8k --points "mode=exact;lanes=1,4" >> gpurun_out/sweep2.jsonl 2>&1
timeout 300 python tools/sweep.py --workload mid --points "mode=exact,fast;lanes=1,4" >> gpurun_out/sweep2.jsonl 2>&1
timeout 300 python tools/sweep.py --workload c32k --points "mode=exact,fast;lanes=1,4" >> gpurun_out/sweep2.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --points "mode=exact,fast;lanes=1,4" >> gpurun_out/sweep2.jsonl 2>&1
cat gpurun_out/sweep2.jsonl
