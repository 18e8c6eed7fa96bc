#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
timeout 300 python tools/sweep.py --workload ns --points "mode=exact;layout=planar,interleaved" > gpurun_out/sweep_sym.jsonl 2>&1
ZG_TUNE_NO_SYM=1 timeout 300 python tools/sweep.py --workload ns --points "mode=exact;layout=planar,interleaved" >> gpurun_out/sweep_sym.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --points "mode=exact;boxes=4;wpc=7;stages=2;late=1,2" >> gpurun_out/sweep_sym.jsonl 2>&1
timeout 300 python tools/sweep.py --workload c32k --points "mode=exact;lanes=1" >> gpurun_out/sweep_sym.jsonl 2>&1
cat gpurun_out/sweep_sym.jsonl
