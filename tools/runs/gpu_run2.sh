#!/bin/bash
# round-1 GPU session 2: tests, tuning sweep, ncu captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest.log
timeout 300 python tools/sweep.py --points "mode=exact,fast;boxes=1,2,4;wpc=0;stages=0" > gpurun_out/sweep1.jsonl 2>&1
timeout 300 python tools/sweep.py --points "mode=exact,fast;boxes=1;wpc=14;stages=2,3" >> gpurun_out/sweep1.jsonl 2>&1
timeout 300 python tools/sweep.py --points "mode=exact,fast;boxes=1,2;wpc=7;stages=2" >> gpurun_out/sweep1.jsonl 2>&1
timeout 300 python tools/sweep.py --points "mode=exact,fast;boxes=1;wpc=7;stages=4" >> gpurun_out/sweep1.jsonl 2>&1
timeout 300 python tools/sweep.py --points "mode=exact,fast;boxes=1,2;layout=interleaved" >> gpurun_out/sweep1.jsonl 2>&1
timeout 300 python tools/sweep.py --points "mode=exact,fast;boxes=1,2;coef=per-channel" >> gpurun_out/sweep1.jsonl 2>&1
timeout 300 python tools/sweep.py --workload c2 --points "mode=exact,fast;boxes=1,2" >> gpurun_out/sweep1.jsonl 2>&1
cat gpurun_out/sweep1.jsonl
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zg_stream_kernel -s 3 -c 1 -o gpurun_out/prof_ns_exact python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_exact.log 2>&1; echo "ncu exact rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zg_stream_kernel -s 3 -c 1 -o gpurun_out/prof_ns_fast python bench.py --steps 2 --warmup 1 --no-e2e --no-cpu --mode fast > gpurun_out/ncu_fast.log 2>&1; echo "ncu fast rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 5 --warmup 3 --no-cpu > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 600 python bench.py > gpurun_out/bench.log 2>&1; echo "bench rc=$?"; cat gpurun_out/bench.log
