#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 300 python tools/sweep.py --workload ns --graph comb --iters 10 --points "mode=exact,fast;layout=planar,interleaved" > gpurun_out/sweep_comb.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --iters 10 --points "mode=exact,fast;layout=interleaved" >> gpurun_out/sweep_comb.jsonl 2>&1
cat gpurun_out/sweep_comb.jsonl
