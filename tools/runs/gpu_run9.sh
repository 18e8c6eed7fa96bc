#!/bin/bash
# early refill A/B on the NS and c3/c5-like shapes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
timeout 300 python tools/sweep.py --workload ns --points "mode=exact,fast;late=0,1" > gpurun_out/sweep_refill.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --points "mode=exact,fast;late=0;stages=3;boxes=1" >> gpurun_out/sweep_refill.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --points "mode=exact;late=0,1;coef=per-channel" >> gpurun_out/sweep_refill.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --points "mode=exact;late=0,1;layout=interleaved" >> gpurun_out/sweep_refill.jsonl 2>&1
timeout 300 python tools/sweep.py --workload c32k --points "mode=exact;late=0,1;lanes=1" >> gpurun_out/sweep_refill.jsonl 2>&1
cat gpurun_out/sweep_refill.jsonl
