#!/bin/bash
# round-1 final checkpoint: tests, smoke, bench lines, launch list of the headline command, ncu of the headline kernel
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench_ns.log 2>&1; echo "bench rc=$?"
timeout 900 python bench.py --workload c2 --no-also > gpurun_out/bench_c2.log 2>&1; echo "bench c2 rc=$?"
timeout 900 python bench.py --mode fast --no-cpu --no-e2e > gpurun_out/bench_ns_fast.log 2>&1; echo "bench fast rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 8 --warmup 3 --no-cpu --no-e2e --no-also > gpurun_out/bench_under_ncu.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:zg_stream_kernel -c 1 -o gpurun_out/ns_exact -f python tools/sweep.py --workload ns --iters 1 --points "mode=exact" > gpurun_out/ncu_ns.log 2>&1; echo "ncu ns rc=$?"
python - <<'PY'
import json
for f in ('bench_ns','bench_c2','bench_ns_fast'):
    d=json.loads(open(f'gpurun_out/{f}.log').read().strip().splitlines()[-1])
    print(f, round(d['value']), round(d['ms_per_step'],4), 'frac', round(d['roofline']['frac'],3), 'e2e', round(d.get('e2e',{}).get('value',0)), 'cpu', round(d.get('cpu_baseline',{}).get('value',0)), d['config']['kernel'], d['clocks']['sm_mhz'], d['clocks']['reasons'])
    for k,v in (d.get('also') or {}).items():
        print('   ', k, round(v.get('value',0)), v.get('ms_per_step'), 'frac', round(v.get('roofline_frac',0),3), v.get('fp32_issue_frac'), v.get('error'))
PY
