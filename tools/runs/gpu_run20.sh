#!/bin/bash
# row pitch: is a power-of-two pitch (32 KB rows) camping on HBM channels?
mkdir -p gpurun_out
rm -f gpurun_out/sweep_pitch.jsonl
for pad in 0 32 64 128 256 1056 2080; do
  timeout 200 python tools/sweep.py --workload ns --pad $pad --iters 10 --points "mode=exact,fast" >> gpurun_out/sweep_pitch.jsonl 2>&1
done
timeout 200 python tools/sweep.py --workload ns --graph copy --pad 0 --iters 10 --points "mode=exact" >> gpurun_out/sweep_pitch.jsonl 2>&1
timeout 200 python tools/sweep.py --workload ns --graph copy --pad 64 --iters 10 --points "mode=exact" >> gpurun_out/sweep_pitch.jsonl 2>&1
timeout 200 python tools/sweep.py --workload ns --graph copy --pad 1056 --iters 10 --points "mode=exact" >> gpurun_out/sweep_pitch.jsonl 2>&1
timeout 200 python tools/sweep.py --workload c2 --pad 64 --iters 10 --points "mode=exact;lanes=4" >> gpurun_out/sweep_pitch.jsonl 2>&1
cat gpurun_out/sweep_pitch.jsonl
