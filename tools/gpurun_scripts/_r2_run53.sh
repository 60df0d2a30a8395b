#!/bin/bash
# full GPU suite + bench after the sampler change (2,048-entry table + compact records when they fit)
timeout 1800 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py > gpurun_out/r2u_bench.json 2> gpurun_out/r2u_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2u_bench.json').read().strip().splitlines()[-1])
print('value %.4g frac %.4f ms %.5f e2e %.4g cfg2d %.4f sampler %.4g (%.4f ms/step) ess/s %.4g' % (d['value'], d['roofline']['frac'], d['ms_per_step'], d['e2e']['value'], d['cfg2_dispersion']['kernel_ms'], d['sampler']['value'], d['sampler']['ms_per_step'], d['ess']['ess_per_sec']))
PY
