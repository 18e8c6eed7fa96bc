#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "fir or long_delay" > gpurun_out/pytest_fir.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_fir.log
timeout 300 python - <<'PY'
import sys, json, torch
sys.path.insert(0, '.')
import zignal_b200 as zg
from zignal_b200 import workloads as wl
C, T = 32768, 8192
g = zg.compile(wl.fir_expr(wl.fir_taps(256)))
for layout, name in ((zg.PLANAR, 'planar'), (zg.INTERLEAVED, 'interleaved')):
    shape = (T, C) if layout == zg.INTERLEAVED else (C, T)
    x = torch.rand(shape, device='cuda') * 2 - 1; y = torch.empty_like(x)
    for mode, mn in ((zg.MODE_EXACT, 'exact'), (zg.MODE_FAST, 'fast')):
        plan = g.plan(channels=C, mode=mode, layout=layout)
        for _ in range(2): plan.process([x], [y])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): plan.process([x], [y])
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(json.dumps({'layout': name, 'mode': mn, 'ms': round(ms, 3), 'msamples': round(C * T / ms / 1e3)}))
PY
