#!/bin/bash
mkdir -p gpurun_out
for v in ng0 ng1; do
  echo "== $v"; NGRTD_LIB=$PWD/build_exp/lib_$v.so python tools/ng_probe2.py 2>&1 | tail -2
done
NGRTD_LIB=$PWD/build_exp/lib_ng1.so timeout 600 python -m pytest tests -m gpu -q -k "ng or sampler or mcmc or noble" 2>&1 | tail -3
for v in ng0 ng1; do
NGRTD_LIB=$PWD/build_exp/lib_$v.so timeout 300 ncu --set full --clock-control none -k regex:k_mcmc_ng -s 2 -c 1 -o gpurun_out/r2g_$v python tools/prof_ng.py > gpurun_out/r2g_ncu_$v.log 2>&1
ncu -i gpurun_out/r2g_$v.ncu-rep --page raw --csv > gpurun_out/r2g_${v}_raw.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2g_${v}_raw.csv > gpurun_out/r2g_${v}_summary.txt 2>/dev/null; echo "== $v"; grep -E "time_duration|long_scoreboard|fp64_cycles|dram__bytes|registers|warps_active" gpurun_out/r2g_${v}_summary.txt
rm -f gpurun_out/r2g_$v.ncu-rep
done
