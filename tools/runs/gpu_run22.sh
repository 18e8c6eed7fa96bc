#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/sweep.py --workload ns --graph copy --iters 10 --points "mode=exact;boxes=8;wpc=3;stages=2;late=1,2" > gpurun_out/sweep_run.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --graph copy --iters 10 --points "mode=exact;boxes=8;wpc=2;stages=3;late=1" >> gpurun_out/sweep_run.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --graph copy --iters 10 --points "mode=exact;boxes=4;wpc=4;stages=3;late=1" >> gpurun_out/sweep_run.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --graph copy --iters 10 --points "mode=exact;boxes=2;wpc=7;stages=4;late=1" >> gpurun_out/sweep_run.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --graph copy --iters 10 --points "mode=exact;boxes=1;wpc=14;stages=4;late=1" >> gpurun_out/sweep_run.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --graph copy --iters 10 --layout interleaved --points "mode=exact;layout=interleaved" >> gpurun_out/sweep_run.jsonl 2>&1
cat gpurun_out/sweep_run.jsonl
