#!/bin/bash
# sampler kernels with 32-bit remainders on the per-step path
python tools/sampler_time.py 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q -x -k "sampler or bmm or mcmc or posterior or trace or population or config or checkpoint or ng_" 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.4g frac %.4f sampler %.4g (%.4f) ess/s %.4g' % (d['value'], d['roofline']['frac'], d['sampler']['value'], d['sampler']['ms_per_step'], d['ess']['ess_per_sec']))"
