#!/bin/bash
# K3 FIR: parity tests, then a first sweep
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k fir > gpurun_out/pytest_fir.log 2>&1; echo "pytest rc=$?"; tail -30 gpurun_out/pytest_fir.log
timeout 600 python tools/sweep_fir.py --points "mode=exact,fast;wpc=4,8,12" > gpurun_out/sweep_fir.jsonl 2>&1
timeout 300 python tools/sweep_fir.py --channels 4096 --samples 65536 --points "mode=exact;wpc=8;segs=0,1" >> gpurun_out/sweep_fir.jsonl 2>&1
cat gpurun_out/sweep_fir.jsonl
