#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -25 > gpurun_out/r2_pytest4.txt
cat gpurun_out/r2_pytest4.txt
NGRTD_LIB=$PWD/build_exp/lib_exp.so timeout 120 python tools/variant_bench.py 2>&1 | tail -1 | tee gpurun_out/r2_variant4.txt
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>gpurun_out/r2_bench4.err | tee gpurun_out/r2_bench4.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['e2e']['sync_call']['value'], 'smp', d['sampler']['value'], d['ess']['ess_per_sec'])"
