#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "long_delay" > gpurun_out/pytest_long.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_long.log
