#!/bin/bash
# round checkpoint: full GPU suite, smoke, bench (both arms), launch list, ncu captures of the headline kernels
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_ns.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_ns.log
timeout 900 python bench.py --workload c2 --no-also > gpurun_out/bench_c2.log 2>&1; echo "bench c2 rc=$?"; tail -1 gpurun_out/bench_c2.log
timeout 900 python bench.py --mode fast --no-cpu --no-e2e > gpurun_out/bench_ns_fast.log 2>&1; echo "bench fast rc=$?"; tail -1 gpurun_out/bench_ns_fast.log
timeout 600 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"; tail -1 gpurun_out/bench_ref.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 4 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zg_stream_kernel -c 1 -o gpurun_out/ns_exact -f python tools/sweep.py --workload ns --iters 1 --points "mode=exact" > gpurun_out/ncu_ns.log 2>&1; echo "ncu ns rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zg_stream_kernel -c 1 -o gpurun_out/ns_fast -f python tools/sweep.py --workload ns --iters 1 --points "mode=fast" > gpurun_out/ncu_ns_fast.log 2>&1; echo "ncu ns fast rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zg_biquad_lanes -c 1 -o gpurun_out/c2_fast -f python tools/sweep.py --workload c2 --iters 1 --points "mode=fast;lanes=4" > gpurun_out/ncu_c2_fast.log 2>&1; echo "ncu c2 fast rc=$?"
