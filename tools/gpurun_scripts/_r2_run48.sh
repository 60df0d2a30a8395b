#!/bin/bash
# full GPU suite + fuzz seeds + bench after the dispersion-tail change (negligible panels skipped, rsqrt)
timeout 1500 python -m pytest tests -m gpu -q -x 2>&1 | tail -3
timeout 600 python tools/fuzz_forward.py 400 21 2>&1 | tail -3
timeout 600 python tools/fuzz_forward.py 400 22 2>&1 | tail -3
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.4g frac %.4f cfg2 %.4f cfg2_disp %.4f cfg5 %.4f sampler %.4g' % (d['value'], d['roofline']['frac'], d['cfg2']['kernel_ms'], d['cfg2_dispersion']['kernel_ms'], d['cfg5']['kernel_ms'], d['sampler']['value']))"
