#!/bin/bash
# re-entry sanity pass: GPU parity tests, smoke, both bench arms
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
timeout 600 python bench.py > gpurun_out/bench_ns.log 2>&1; echo "bench rc=$?"; tail -3 gpurun_out/bench_ns.log
timeout 600 python bench.py --impl reference > gpurun_out/bench_ref.log 2>&1; echo "bench ref rc=$?"; tail -3 gpurun_out/bench_ref.log
nproc; nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm,power.limit --format=csv
