#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "cuda_graph" > gpurun_out/pytest_graph.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_graph.log
