#!/bin/bash
mkdir -p gpurun_out
L=$PWD/build_exp
{
NGRTD_LIB=$L/lib_r1.so timeout 120 python tools/r2_probe.py
NGRTD_LIB=$L/lib_exp.so timeout 120 python tools/r2_probe.py
NGRTD_TAPE_MIN=100000 NGRTD_LIB=$L/lib_exp.so timeout 120 python tools/r2_probe.py
NGRTD_TAPE_MIN=32 NGRTD_LIB=$L/lib_exp.so timeout 120 python tools/r2_probe.py
NGRTD_PDL=0 NGRTD_LIB=$L/lib_exp.so timeout 120 python tools/r2_probe.py
} > gpurun_out/r2_probe2.txt 2>&1
cat gpurun_out/r2_probe2.txt
NGRTD_LIB=$L/lib_exp.so timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_forward -s 6 -c 1 -o gpurun_out/r2_fwd_tape python tools/prof_one.py 8 > gpurun_out/r2_ncu2.log 2>&1
tail -3 gpurun_out/r2_ncu2.log
ncu -i gpurun_out/r2_fwd_tape.ncu-rep --page raw --csv > gpurun_out/r2_fwd_tape_raw.csv 2>/dev/null
ncu -i gpurun_out/r2_fwd_tape.ncu-rep --page source --csv > gpurun_out/r2_fwd_tape_src.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2_fwd_tape_raw.csv gpurun_out/r2_fwd_tape_src.csv | head -60
