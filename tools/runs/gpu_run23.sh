#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/sweep.py --workload ns --iters 10 --points "mode=exact,fast;lanes=4;boxes=8,16" > gpurun_out/sweep_ns_lanes.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --iters 10 --points "mode=exact;lanes=4;boxes=8;wpc=6,8" >> gpurun_out/sweep_ns_lanes.jsonl 2>&1
timeout 300 python tools/sweep.py --workload c32k --iters 10 --points "mode=exact,fast;lanes=4,1" >> gpurun_out/sweep_ns_lanes.jsonl 2>&1
timeout 300 python tools/sweep.py --workload mid --iters 10 --points "mode=exact,fast;lanes=4,1" >> gpurun_out/sweep_ns_lanes.jsonl 2>&1
cat gpurun_out/sweep_ns_lanes.jsonl
