#!/bin/bash
# L2 bulk prefetch of the input rows ahead of the TMA boxes: window x distance, on the headline kernel and two neighbours
mkdir -p gpurun_out
: > gpurun_out/sweep_pf.jsonl
timeout 300 python tools/sweep.py --workload ns --iters 20 --points "mode=exact;pf=0,1024,2048,4096;pfd=2,4" >> gpurun_out/sweep_pf.jsonl 2>&1; echo "rc=$?"
timeout 300 python tools/sweep.py --workload ns --iters 20 --points "mode=exact;pf=1024,2048;pfd=2;boxes=2;wpc=14" >> gpurun_out/sweep_pf.jsonl 2>&1; echo "rc=$?"
timeout 300 python tools/sweep.py --workload ns --iters 20 --points "mode=fast;pf=0,1024,2048,4096;pfd=2" >> gpurun_out/sweep_pf.jsonl 2>&1; echo "rc=$?"
timeout 300 python tools/sweep.py --workload ns --iters 20 --graph copy --points "mode=exact;pf=0,1024,2048,4096;pfd=2" >> gpurun_out/sweep_pf.jsonl 2>&1; echo "rc=$?"
cut -c1-330 gpurun_out/sweep_pf.jsonl
