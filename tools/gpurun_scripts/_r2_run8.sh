#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_pytest8.txt
cat gpurun_out/r2_pytest8.txt
timeout 300 python tools/dm_tail_check.py 2>&1 | tail -8 | tee gpurun_out/r2_dmtail8.txt
timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/r2_bench8.err > gpurun_out/r2_bench8.json
tail -5 gpurun_out/r2_bench8.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench8.json'))
print('value',d['value'],'frac',d['roofline']['frac'],'kms',d['roofline']['kernel_ms'],'iso',d['roofline']['kernel_ms_isolated'],'nan',d['nan_frac'])
print('e2e',d['e2e']['value'],'sync',d['e2e']['sync_call']['value'],'link',d['e2e']['host_link'])
for k in ('informative','cfg2','cfg2_dispersion','cfg5','sampler','ess','timing','clocks','cpu_baseline'):
    print(k, d.get(k))
PY
