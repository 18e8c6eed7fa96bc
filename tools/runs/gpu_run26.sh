#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 300 python tools/sweep.py --workload ns --iters 20 --points "mode=exact;layout=planar,interleaved" > gpurun_out/sweep_sym2.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --iters 20 --points "mode=exact;wpc=14;boxes=2" >> gpurun_out/sweep_sym2.jsonl 2>&1
timeout 300 python tools/sweep.py --workload c32k --iters 20 --points "mode=exact;lanes=1" >> gpurun_out/sweep_sym2.jsonl 2>&1
timeout 300 python tools/sweep.py --workload c32k --iters 20 --points "mode=exact;lanes=1;wpc=14;boxes=2" >> gpurun_out/sweep_sym2.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --sections 2 --iters 20 --points "mode=exact" >> gpurun_out/sweep_sym2.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --sections 8 --iters 20 --points "mode=exact" >> gpurun_out/sweep_sym2.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --sections 8 --iters 20 --points "mode=exact;wpc=14;boxes=2" >> gpurun_out/sweep_sym2.jsonl 2>&1
cat gpurun_out/sweep_sym2.jsonl
