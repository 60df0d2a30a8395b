python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/s2_tests.txt
cat gpurun_out/s2_tests.txt
python bench.py > gpurun_out/bench_s2_b.json 2> gpurun_out/bench_s2_b.err; tail -c 1500 gpurun_out/bench_s2_b.err; python -c "
import json; d=json.load(open('gpurun_out/bench_s2_b.json')); print(d['value'], d['e2e'], d['roofline']['frac'], d['sampler']['value'])"
