#!/bin/bash
# ncu capture of the sampler kernel (new instantiation: 2,048-entry table, compact records)
timeout 600 ncu --set full --import-source on --clock-control none -k regex:k_mcmc_age -s 2 -c 1 -o gpurun_out/r2v_age python tools/prof_sampler.py 8 > gpurun_out/r2v_ncu_age.log 2>&1
ncu -i gpurun_out/r2v_age.ncu-rep --page raw --csv > gpurun_out/r2v_age_raw.csv 2>/dev/null
ncu -i gpurun_out/r2v_age.ncu-rep --page source --csv > gpurun_out/r2v_age_src.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2v_age_raw.csv gpurun_out/r2v_age_src.csv > gpurun_out/r2v_age_summary.txt 2>/dev/null; head -40 gpurun_out/r2v_age_summary.txt
rm -f gpurun_out/r2v_age.ncu-rep
