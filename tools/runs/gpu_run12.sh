#!/bin/bash
# K1b: branch-free fill/drain iterations + per-lane address tables
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
timeout 300 python tools/sweep.py --workload c2 --points "mode=exact,fast;lanes=4" > gpurun_out/sweep_k1b.jsonl 2>&1
timeout 300 python tools/sweep.py --workload c2 --points "mode=exact;lanes=4;boxes=8,32;stages=2,3" >> gpurun_out/sweep_k1b.jsonl 2>&1
timeout 300 python tools/sweep.py --workload c8k --points "mode=exact;lanes=4,1" >> gpurun_out/sweep_k1b.jsonl 2>&1
cat gpurun_out/sweep_k1b.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zg_biquad_lanes -c 1 -o gpurun_out/k1b_exact -f python tools/sweep.py --workload c2 --iters 1 --points "mode=exact;lanes=4" > gpurun_out/ncu_k1b.log 2>&1; echo "ncu rc=$?"
