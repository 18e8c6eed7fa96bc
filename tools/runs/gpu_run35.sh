#!/bin/bash
# last verification of round 1: the three things the driver runs at round end
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_ns.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_ns.log | cut -c1-1200
