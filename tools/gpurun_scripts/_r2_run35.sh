#!/bin/bash
mkdir -p gpurun_out
for l in noblegas_rtd_mcmc_b200/libngrtd.so build_exp/lib_exp.so noblegas_rtd_mcmc_b200/libngrtd.so build_exp/lib_exp.so; do
NGRTD_LIB=$PWD/$l timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null > gpurun_out/r2q.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2q.json'))
print('$l value %.4g e2e %.4g %.4f sync %.4g' % (d['value'],d['e2e']['value'],d['e2e']['ms_per_step'],d['e2e']['sync_call']['value']))
PY
done
NGRTD_LIB=$PWD/build_exp/lib_exp.so timeout 600 python -m pytest tests/test_fullsize_gpu.py tests/test_r2_gpu.py -m gpu -q -k "host or submit or copy or pinned" 2>&1 | tail -3
