#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_pytest9.txt
cat gpurun_out/r2_pytest9.txt
timeout 900 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>gpurun_out/r2_bench9.err > gpurun_out/r2_bench9.json
tail -3 gpurun_out/r2_bench9.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench9.json'))
print('value',d['value'],'frac',d['roofline']['frac'],'kms',d['roofline']['kernel_ms'],'iso',d['roofline']['kernel_ms_isolated'],'nan',d['nan_frac'])
print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'sync',d['e2e']['sync_call']['value'],'link',d['e2e']['host_link']['gbps_per_rank'], d['e2e']['host_link']['e2e_ceiling_evals_per_s'])
for k in ('informative','cfg2','cfg2_dispersion','cfg5','sampler','ess'):
    print(k, {kk:vv for kk,vv in d.get(k).items() if kk in ('kernel_ms','ms_per_step','value','frac','ess_per_sec','seconds','nan_frac','L_eff')})
PY
