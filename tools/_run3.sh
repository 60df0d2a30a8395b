python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python -c "import __graft_entry__ as g; g.smoke()"
for i in 1 2 3; do
python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>/dev/null | tee gpurun_out/s3_bench_$i.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['e2e']['sync_call']['value'], 'smp', d['sampler']['value'], d['ess']['ess_per_sec'])"
done
