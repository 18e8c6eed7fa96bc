#!/bin/bash
# checkpoint after the front-end additions (ResultType, spellings), the StreamArgs change (L2 prefetch knob) and the
# peer-mapping edge step: GPU tests, smoke, the default bench line, the reference arm
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_ns.log 2>&1; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref_short.log 2>&1; echo "ref rc=$?"
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_ns.log').read().strip().splitlines()[-1])
print('ns', round(d['value']), round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d.get('e2e',{}).get('value',0)), 'cpu', round(d.get('cpu_baseline',{}).get('value',0)), d['config']['kernel'], d['clocks']['sm_mhz'], d['clocks']['reasons'], d['cpu_baseline'].get('parity_spot_check'))
for k,v in (d.get('also') or {}).items():
    print('   ', k, round(v.get('value',0)), v.get('ms_per_step'), 'frac', round(v.get('roofline_frac',0),3), v.get('fp32_issue_frac'), v.get('error'))
r=json.loads(open('gpurun_out/bench_ref_short.log').read().strip().splitlines()[-1]); print('ref', round(r['value']), r['cpu_baseline']['cores'])
PY
