#!/bin/bash
mkdir -p gpurun_out
python -c "
from noblegas_rtd_mcmc_b200 import _lib
for i in range(3): print(_lib.fp64_peak_probe())
"
timeout 600 python -m pytest tests/test_r2_gpu.py -m gpu -q 2>&1 | tail -3
timeout 600 python bench.py --steps 20 --warmup 5 --no-extras --no-cpu-baseline 2>/dev/null > gpurun_out/r2n.json
python - <<PY
import json
d=json.load(open('gpurun_out/r2n.json'))
print(d['roofline']['frac'], d['roofline']['peak_live'], d['e2e']['value'])
PY
