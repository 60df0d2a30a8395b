#!/bin/bash
for i in 1 2; do timeout 600 python bench.py --no-cpu-baseline --no-extras 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('value %.4g frac %.4f ms %.5f e2e %.4g' % (d['value'], d['roofline']['frac'], d['ms_per_step'], d['e2e']['value']))"; done
python tools/variant_bench.py 2>&1 | tail -1
