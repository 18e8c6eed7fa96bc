#!/bin/bash
# memory-side knobs on NS; source-level profile of K1b
mkdir -p gpurun_out
timeout 300 python tools/sweep.py --workload ns --points "mode=fast,exact;hint=0,1,2,3" > gpurun_out/sweep_mem.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --points "mode=fast,exact;promo=1,4" >> gpurun_out/sweep_mem.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --points "mode=fast;late=1,2;wpc=7;boxes=4;stages=2;hint=0,3" >> gpurun_out/sweep_mem.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --points "mode=fast;late=1;wpc=5;boxes=5;stages=2" >> gpurun_out/sweep_mem.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --points "mode=fast;late=1;wpc=4;boxes=6;stages=2" >> gpurun_out/sweep_mem.jsonl 2>&1
timeout 300 python tools/sweep.py --workload c2 --points "mode=exact;lanes=4" >> gpurun_out/sweep_mem.jsonl 2>&1
cat gpurun_out/sweep_mem.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zg_biquad_lanes -c 1 -o gpurun_out/k1b_exact -f python tools/sweep.py --workload c2 --iters 1 --points "mode=exact;lanes=4" > gpurun_out/ncu_k1b.log 2>&1; echo "ncu rc=$?"
