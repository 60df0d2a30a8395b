python -m pytest tests -m gpu -x -q 2>&1 | tail -6
python bench.py --steps 200 --warmup 20 > gpurun_out/s2_bench_c.json 2> gpurun_out/s2_bench_c.err; tail -5 gpurun_out/s2_bench_c.err; python -c "
import json; d=json.load(open('gpurun_out/s2_bench_c.json')); print(d['value'], d['roofline']['frac'], json.dumps(d['e2e'], indent=1), d['ess'])"
