#!/bin/bash
# random graphs on the device against the oracle
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_fuzz_gpu.py -m gpu -q -x > gpurun_out/pytest_fuzz.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_fuzz.log | cut -c1-400
