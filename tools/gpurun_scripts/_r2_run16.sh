#!/bin/bash
mkdir -p gpurun_out
NGRTD_LIB=$PWD/build_exp/lib_exp.so timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_forward -s 6 -c 1 -o gpurun_out/r2_fwd_c python tools/prof_one.py 8 > gpurun_out/r2_ncu16.log 2>&1
ncu -i gpurun_out/r2_fwd_c.ncu-rep --page raw --csv > gpurun_out/r2_fwd_c_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_fwd_c.ncu-rep --page source --csv > gpurun_out/r2_fwd_c_src.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2_fwd_c_raw.csv 2>/dev/null | head -40
