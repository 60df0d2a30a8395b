#!/bin/bash
mkdir -p gpurun_out
v=ng1
NGRTD_LIB=$PWD/build_exp/lib_$v.so timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_mcmc_ng -s 2 -c 1 -o gpurun_out/r2g_$v python tools/prof_ng.py > gpurun_out/r2g_ncu_$v.log 2>&1
ncu -i gpurun_out/r2g_$v.ncu-rep --page raw --csv > gpurun_out/r2g_${v}_raw.csv 2>/dev/null
ncu -i gpurun_out/r2g_$v.ncu-rep --page source --csv > gpurun_out/r2g_${v}_src.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2g_${v}_raw.csv gpurun_out/r2g_${v}_src.csv > gpurun_out/r2g_${v}_summary.txt 2>/dev/null; tail -40 gpurun_out/r2g_${v}_summary.txt
rm -f gpurun_out/r2g_$v.ncu-rep
