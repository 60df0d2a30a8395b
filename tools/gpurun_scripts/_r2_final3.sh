#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -6 > gpurun_out/r2p_pytest.txt; cat gpurun_out/r2p_pytest.txt
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 2>gpurun_out/r2p_bench.err > gpurun_out/r2p_bench.json; tail -2 gpurun_out/r2p_bench.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2p_bench_ref.json 2>/dev/null
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2p_bench.json'))
print('value',d['value'],'frac',d['roofline']['frac'],'kms',d['roofline']['kernel_ms'],'iso',d['roofline']['kernel_ms_isolated'],'nan',d['nan_frac'])
print('e2e',d['e2e']['value'],d['e2e']['ms_per_step'],'sync',d['e2e']['sync_call']['value'],'link',d['e2e']['host_link']['gbps_per_rank'])
for k in ('informative','cfg2','cfg2_dispersion','cfg5','sampler','ess','cpu_baseline'):
    print(k, {kk:vv for kk,vv in d.get(k).items() if kk in ('kernel_ms','ms_per_step','value','frac','ess_per_sec','seconds','nan_frac','L_eff','cores')})
PY
# ncu launch list of the bench command (cold-cache, serialised: shares only)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2p_launches.csv python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-extras > gpurun_out/r2p_ncu_bench.log 2>&1
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_forward -s 6 -c 1 -o gpurun_out/r2p_fwd python tools/prof_one.py 8 > gpurun_out/r2p_ncu_fwd.log 2>&1
ncu -i gpurun_out/r2p_fwd.ncu-rep --page raw --csv > gpurun_out/r2p_fwd_raw.csv 2>/dev/null
ncu -i gpurun_out/r2p_fwd.ncu-rep --page source --csv > gpurun_out/r2p_fwd_src.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2p_fwd_raw.csv gpurun_out/r2p_fwd_src.csv > gpurun_out/r2p_fwd_summary.txt 2>/dev/null; head -12 gpurun_out/r2p_fwd_summary.txt
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_mcmc_ng -s 2 -c 1 -o gpurun_out/r2p_ng python tools/prof_ng.py > gpurun_out/r2p_ncu_ng.log 2>&1
ncu -i gpurun_out/r2p_ng.ncu-rep --page raw --csv > gpurun_out/r2p_ng_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2p_ng_raw.csv > gpurun_out/r2p_ng_summary.txt 2>/dev/null; grep -E "time_duration|long_scoreboard|fp64_cycles|dram__bytes" gpurun_out/r2p_ng_summary.txt
rm -f gpurun_out/r2p_fwd.ncu-rep gpurun_out/r2p_ng.ncu-rep

