python tools/ng_probe2.py
for s in 0 1 0 1; do NGRTD_STAGE=$s python bench.py --steps 100 --warmup 10 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print('stage $s', d['value'], d['ms_per_step'], d['roofline']['kernel_ms'], d['roofline']['frac'], d['e2e']['value'])"; done
