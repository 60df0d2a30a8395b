#!/bin/bash
mkdir -p gpurun_out
for wc in 1 0 1 0 0; do
NGRTD_BENCH_WC=$wc timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null > gpurun_out/r2k.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2k.json'))
print('wc$wc value %.4g e2e %.4g %.4f sync %.4g h2d %.4f wc %.4f duplex %.4f' % (d['value'],d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e']['sync_call']['value'], d['e2e']['host_link']['h2d_ms_per_batch'], d['e2e']['host_link']['h2d_ms_per_batch_write_combined'], d['e2e']['host_link']['duplex_ms_per_batch']))
PY
done
