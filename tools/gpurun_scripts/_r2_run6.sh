#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -5 > gpurun_out/r2_pytest6.txt
cat gpurun_out/r2_pytest6.txt
timeout 600 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>gpurun_out/r2_bench6.err | tee gpurun_out/r2_bench6.json | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['roofline']['frac'], d['roofline']['kernel_ms'], 'e2e', d['e2e']['value'], d['e2e']['sync_call']['value'], 'smp', d['sampler']['value'], d['ess']['ess_per_sec'])"
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_forward -s 6 -c 1 -o gpurun_out/r2_fwd_b python tools/prof_one.py 8 > gpurun_out/r2_ncu6.log 2>&1
ncu -i gpurun_out/r2_fwd_b.ncu-rep --page raw --csv > gpurun_out/r2_fwd_b_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_fwd_b.ncu-rep --page source --csv > gpurun_out/r2_fwd_b_src.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2_fwd_b_raw.csv gpurun_out/r2_fwd_b_src.csv 2>/dev/null | head -45
