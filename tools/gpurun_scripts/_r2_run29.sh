#!/bin/bash
mkdir -p gpurun_out
for wc in 0 1; do
NGRTD_BENCH_WC=$wc timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>gpurun_out/r2j_bench_wc$wc.err > gpurun_out/r2j_bench_wc$wc.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2j_bench_wc$wc.json'))
print('wc$wc value',d['value'],'e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'sync',d['e2e']['sync_call']['value'])
print({k:v for k,v in d['e2e']['host_link'].items() if k!='note'})
PY
done
