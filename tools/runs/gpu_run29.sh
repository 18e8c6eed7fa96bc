#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "on_the_device or per_channel" > gpurun_out/pytest_new.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_new.log
