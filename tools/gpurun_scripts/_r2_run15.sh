#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_r2_gpu.py tests/test_bmm_posterior_gpu.py -q 2>&1 | tail -6
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 2>gpurun_out/r2_bench_n2.err > gpurun_out/r2_bench_n2.json
tail -3 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n2.json'))
print('N',d['n_gpus'],'value',d['value'],'frac',d['roofline']['frac'],'kms',d['roofline']['kernel_ms'],'per_rank',d['per_rank_ms_per_step'])
print('e2e',d['e2e']['value'],'link',d['e2e']['host_link']['gbps_per_rank'],d['e2e']['host_link']['gbps_aggregate'])
for k in ('cfg2','cfg5','sampler','ess'):
    print(k, {kk:vv for kk,vv in d.get(k).items() if kk in ('kernel_ms','ms_per_step','value','frac','ess_per_sec','seconds','chains')})
PY
timeout 300 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 | cut -c1-400
