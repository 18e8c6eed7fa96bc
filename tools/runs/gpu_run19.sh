#!/bin/bash
# write-only graphs (configs[2], configs[4]) and the skeleton's own ceiling (copy graph): geometry
mkdir -p gpurun_out
timeout 300 python tools/sweep.py --graph osc --workload c3 --points "mode=exact;boxes=1,2,4;stages=2,3" > gpurun_out/sweep_synth.jsonl 2>&1
timeout 300 python tools/sweep.py --graph osc --workload c3 --points "mode=exact;boxes=4,8;stages=2;wpc=7" >> gpurun_out/sweep_synth.jsonl 2>&1
timeout 300 python tools/sweep.py --graph osc --workload c3 --points "mode=exact;boxes=2;stages=4;wpc=7" >> gpurun_out/sweep_synth.jsonl 2>&1
timeout 300 python tools/sweep.py --graph poly --workload c5 --points "mode=exact,fast;boxes=2,4;stages=2" >> gpurun_out/sweep_synth.jsonl 2>&1
timeout 300 python tools/sweep.py --graph copy --workload ns --points "mode=exact;boxes=2;stages=2;late=1,2" >> gpurun_out/sweep_synth.jsonl 2>&1
timeout 300 python tools/sweep.py --graph copy --workload ns --points "mode=exact;boxes=4;stages=2;wpc=7;late=1,2" >> gpurun_out/sweep_synth.jsonl 2>&1
cat gpurun_out/sweep_synth.jsonl
