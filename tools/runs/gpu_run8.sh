#!/bin/bash
# bf16 storage + FIR: full GPU suite, bench with all configs, ncu capture of the FIR kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest.log
timeout 900 python bench.py --no-cpu > gpurun_out/bench_ns.log 2>&1; echo "bench rc=$?"; tail -2 gpurun_out/bench_ns.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zg_fir -c 1 -o gpurun_out/fir_exact -f python tools/sweep_fir.py --iters 1 --points "mode=exact;wpc=12" > gpurun_out/ncu_fir_exact.log 2>&1; echo "ncu rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zg_fir -c 1 -o gpurun_out/fir_fast -f python tools/sweep_fir.py --iters 1 --points "mode=fast;wpc=12" > gpurun_out/ncu_fir_fast.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out
