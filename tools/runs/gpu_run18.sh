#!/bin/bash
mkdir -p gpurun_out
timeout 120 python tools/membw.py | tee gpurun_out/membw.json
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 900 python bench.py --no-e2e > gpurun_out/bench_ns.log 2>&1; echo "bench rc=$?"
timeout 900 python bench.py --no-e2e --no-cpu --mode fast > gpurun_out/bench_ns_fast.log 2>&1; echo "bench rc=$?"
python - <<'PY'
import json
for f in ('bench_ns','bench_ns_fast'):
    d=json.loads(open(f'gpurun_out/{f}.log').read().strip().splitlines()[-1])
    print(f, round(d['value']), round(d['roofline']['frac'],3), d.get('cpu_baseline',{}).get('parity_spot_check'))
    for k,v in (d.get('also') or {}).items():
        print('   ', k, round(v.get('value',0)), v.get('ms_per_step'), 'frac', round(v.get('roofline_frac',0),3), v.get('fp32_issue_frac'), v.get('error'))
PY
