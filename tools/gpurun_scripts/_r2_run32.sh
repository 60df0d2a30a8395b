#!/bin/bash
mkdir -p gpurun_out
timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_mcmc_age -s 2 -c 1 -o gpurun_out/r2m_age python tools/prof_sampler.py 20 > gpurun_out/r2m_ncu_age.log 2>&1
ncu -i gpurun_out/r2m_age.ncu-rep --page raw --csv > gpurun_out/r2m_age_raw.csv 2>/dev/null
ncu -i gpurun_out/r2m_age.ncu-rep --page source --csv > gpurun_out/r2m_age_src.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2m_age_raw.csv gpurun_out/r2m_age_src.csv > gpurun_out/r2m_age_summary.txt 2>/dev/null; head -60 gpurun_out/r2m_age_summary.txt
rm -f gpurun_out/r2m_age.ncu-rep
