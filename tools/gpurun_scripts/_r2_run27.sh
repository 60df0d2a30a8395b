#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -4
timeout 900 python bench.py --steps 20 --warmup 5 --no-extras 2>gpurun_out/r2i_bench.err > gpurun_out/r2i_bench.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2i_bench.json'))
print('value',d['value'],'frac',d['roofline']['frac'],'kms',d['roofline']['kernel_ms'],'iso',d['roofline']['kernel_ms_isolated'],'nan',d['nan_frac'])
print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'sync',d['e2e']['sync_call']['value'])
print(d.get('informative'))
PY
