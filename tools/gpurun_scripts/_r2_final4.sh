#!/bin/bash
# session-3 final captures on one B200: GPU suite, smoke, bench (own arm + reference arm), ncu launch list of the bench command
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -4 > gpurun_out/r2w_pytest.txt; cat gpurun_out/r2w_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 900 python bench.py 2>gpurun_out/r2w_bench.err > gpurun_out/r2w_bench.json; tail -2 gpurun_out/r2w_bench.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2w_bench_ref.json 2>/dev/null
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2w_bench.json').read().strip().splitlines()[-1])
print('value',d['value'],'frac',d['roofline']['frac'],'kms',d['roofline']['kernel_ms'],'iso',d['roofline']['kernel_ms_isolated'],'nan',d['nan_frac'])
print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'sync',d['e2e']['sync_call']['value'],'link',d['e2e']['host_link']['gbps_per_rank'])
for k in ('informative','cfg2','cfg2_dispersion','cfg5','sampler','ess','cpu_baseline'):
    print(k, {kk:vv for kk,vv in d.get(k).items() if kk in ('kernel_ms','ms_per_step','value','frac','ess_per_sec','seconds','nan_frac','L_eff','cores')})
r=json.loads(open('gpurun_out/r2w_bench_ref.json').read().strip().splitlines()[-1])
print('reference arm', r.get('value'), r.get('cpu_baseline'))
PY
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2w_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2w_ncu_bench.log 2>&1
grep -c k_forward gpurun_out/r2w_launches.csv
