#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
timeout 300 python tools/sweep.py --workload c2 --points "mode=exact,fast;lanes=4" > gpurun_out/sweep4.jsonl 2>&1
ZG_TUNE_NO3D=1 timeout 300 python tools/sweep.py --workload c2 --points "mode=exact;lanes=4" >> gpurun_out/sweep4.jsonl 2>&1
timeout 300 python tools/sweep.py --workload c2 --points "mode=exact;lanes=4;boxes=8,16;stages=2,3,4" >> gpurun_out/sweep4.jsonl 2>&1
timeout 300 python tools/sweep.py --workload c8k --points "mode=exact;lanes=1,4" >> gpurun_out/sweep4.jsonl 2>&1
timeout 300 python tools/sweep.py --workload mid --points "mode=exact,fast;lanes=1,4" >> gpurun_out/sweep4.jsonl 2>&1
timeout 300 python tools/sweep.py --workload c32k --points "mode=exact;lanes=1,4" >> gpurun_out/sweep4.jsonl 2>&1
cat gpurun_out/sweep4.jsonl
timeout 600 python bench.py --no-cpu > gpurun_out/bench_ns.log 2>&1; echo "bench rc=$?"; cat gpurun_out/bench_ns.log
timeout 600 python bench.py --no-cpu --workload c2 > gpurun_out/bench_c2.log 2>&1; echo "bench rc=$?"; cat gpurun_out/bench_c2.log
