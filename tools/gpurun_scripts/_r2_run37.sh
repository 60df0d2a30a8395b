#!/bin/bash
# full GPU suite + smoke + bench on the committed state (session 3 baseline)
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2s_pytest.txt
cat gpurun_out/r2s_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/r2s_bench.json 2> gpurun_out/r2s_bench.err
python - <<'PY'
import json
d=json.loads(open('gpurun_out/r2s_bench.json').read().strip().splitlines()[-1])
print('value %.4g frac %.4f ms %.5f e2e %.4g' % (d['value'], d['roofline']['frac'], d['ms_per_step'], d['e2e']['value']))
PY
