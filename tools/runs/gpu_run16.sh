#!/bin/bash
# chunk unroll of the streaming skeleton (instruction-cache footprint); reference arm thread count under OMP_NUM_THREADS=1
mkdir -p gpurun_out
timeout 300 python tools/sweep.py --workload ns --points "mode=exact;jit=1;cu=8,4,2" > gpurun_out/sweep_cu.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --points "mode=fast;jit=1;cu=8,4" >> gpurun_out/sweep_cu.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --points "mode=exact,fast;jit=0" >> gpurun_out/sweep_cu.jsonl 2>&1
timeout 300 python tools/sweep.py --workload ns --points "mode=exact;jit=1;cu=8,4;coef=per-channel" >> gpurun_out/sweep_cu.jsonl 2>&1
cat gpurun_out/sweep_cu.jsonl
OMP_NUM_THREADS=1 timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_omp1.log 2>&1; tail -1 gpurun_out/bench_ref_omp1.log | cut -c1-200
