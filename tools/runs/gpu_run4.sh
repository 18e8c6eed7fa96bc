#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
timeout 300 python tools/sweep.py --workload c2 --points "mode=exact,fast;lanes=1,4" > gpurun_out/sweep3.jsonl 2>&1
timeout 300 python tools/sweep.py --workload c2 --points "mode=exact;lanes=4;boxes=8,32;stages=2,3,4" >> gpurun_out/sweep3.jsonl 2>&1
timeout 300 python tools/sweep.py --workload c8k --points "mode=exact;lanes=1,4" >> gpurun_out/sweep3.jsonl 2>&1
timeout 300 python tools/sweep.py --workload mid --points "mode=exact,fast;lanes=1,4" >> gpurun_out/sweep3.jsonl 2>&1
timeout 300 python tools/sweep.py --workload c32k --points "mode=exact,fast;lanes=1,4" >> gpurun_out/sweep3.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --points "mode=exact,fast;lanes=1" >> gpurun_out/sweep3.jsonl 2>&1
cat gpurun_out/sweep3.jsonl
