#!/bin/bash
# K1b continuous pipeline
mkdir -p gpurun_out
timeout 240 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "lanes or config2" > gpurun_out/pytest_lanes.log 2>&1; rc=$?; echo "pytest lanes rc=$rc"; tail -15 gpurun_out/pytest_lanes.log
if [ $rc -ne 0 ]; then exit 1; fi
timeout 300 python tools/sweep.py --workload c2 --points "mode=exact,fast;lanes=4" > gpurun_out/sweep_k1b.jsonl 2>&1
timeout 300 python tools/sweep.py --workload c2 --points "mode=exact;lanes=4;boxes=8,32;stages=2,3" >> gpurun_out/sweep_k1b.jsonl 2>&1
timeout 300 python tools/sweep.py --workload c8k --points "mode=exact;lanes=4,1" >> gpurun_out/sweep_k1b.jsonl 2>&1
cat gpurun_out/sweep_k1b.jsonl
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zg_biquad_lanes -c 1 -o gpurun_out/k1b_exact -f python tools/sweep.py --workload c2 --iters 1 --points "mode=exact;lanes=4" > gpurun_out/ncu_k1b.log 2>&1; echo "ncu rc=$?"
