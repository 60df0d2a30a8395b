#!/bin/bash
# dispersion tail: negligible panels skipped, rsqrt -- parity (goldens, fuzz, extreme parameters) and the bench blocks
timeout 900 python -m pytest tests -m gpu -q -x -k "tail or dm_ or extreme or golden or edges or fuzz or real" 2>&1 | tail -3
timeout 600 python tools/fuzz_forward.py 300 11 2>&1 | tail -4
timeout 300 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.4g frac %.4f cfg2 %.4f cfg2_disp %.4f cfg5 %.4f sampler %.4g' % (d['value'], d['roofline']['frac'], d['cfg2']['kernel_ms'], d['cfg2_dispersion']['kernel_ms'], d['cfg5']['kernel_ms'], d['sampler']['value']))"
