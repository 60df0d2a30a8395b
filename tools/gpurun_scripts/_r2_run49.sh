#!/bin/bash
# 2-GPU bench line (torchrun, NCCL) + the multi-GPU GPU tests on the final session-3 build
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 > gpurun_out/r2t_bench_n2.json 2> gpurun_out/r2t_bench_n2.err
tail -c 600 gpurun_out/r2t_bench_n2.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2t_bench_n2.json').read().strip().splitlines()[-1])
print('N=2 value %.4g frac %.4f e2e %.4g ess/s %.4g' % (d['value'], d['roofline']['frac'], d['e2e']['value'], d['ess']['ess_per_sec']))
PY
timeout 600 python -m pytest tests -m gpu -q -x -k "two or multi or shard or device" 2>&1 | tail -3
