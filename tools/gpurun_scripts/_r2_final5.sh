#!/bin/bash
# multi-GPU bench lines (torchrun, NCCL), N taken from the first argument
N=${1:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N > gpurun_out/r2w_bench_n$N.json 2> gpurun_out/r2w_bench_n$N.err
python - $N <<'PY'
import json, sys
n=sys.argv[1]
d=json.loads(open('gpurun_out/r2w_bench_n%s.json' % n).read().strip().splitlines()[-1])
print('N=%s value %.4g frac %.4f e2e %.4g link %.1f GB/s/rank sampler %.4g ess/s %.4g' % (n, d['value'], d['roofline']['frac'], d['e2e']['value'], d['e2e']['host_link']['gbps_per_rank'], d['sampler']['value'], d['ess']['ess_per_sec']))
PY
