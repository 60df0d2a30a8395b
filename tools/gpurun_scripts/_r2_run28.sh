#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 2>gpurun_out/r2_bench_n8.err > gpurun_out/r2_bench_n8.json
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2_bench_n8.json'))
print('value',d['value'],'frac',d['roofline']['frac'],'kms',d['roofline']['kernel_ms'])
print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'link',d['e2e']['host_link'])
for k in ('cfg2','cfg5','sampler','ess'):
    print(k, {kk:vv for kk,vv in d.get(k).items() if kk in ('kernel_ms','ms_per_step','value','frac','ess_per_sec','seconds')})
PY
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29514 examples/config4_joint_fit.py 1048576 256 > gpurun_out/r2_config4_1M_8gpu.txt 2>gpurun_out/r2_config4_1M_8gpu.err; head -8 gpurun_out/r2_config4_1M_8gpu.txt; tail -3 gpurun_out/r2_config4_1M_8gpu.err
