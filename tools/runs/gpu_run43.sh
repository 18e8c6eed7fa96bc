#!/bin/bash
# compute-sanitizer memcheck over smoke(): every kernel family once (K1, K1b, generated + NVRTC, FIR, bf16 storage)
mkdir -p gpurun_out
timeout 110 compute-sanitizer --tool memcheck --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/memcheck_smoke.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|smoke ok|Invalid|out of bounds|misaligned" gpurun_out/memcheck_smoke.log | head -10; tail -3 gpurun_out/memcheck_smoke.log | cut -c1-200
