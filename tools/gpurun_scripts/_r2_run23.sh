#!/bin/bash
mkdir -p gpurun_out
for v in ngr7 ngr8; do
  echo "== $v"; NGRTD_LIB=$PWD/build_exp/lib_$v.so python tools/ng_probe2.py 2>&1 | tail -1
  NGRTD_LIB=$PWD/build_exp/lib_$v.so python tools/ng_probe2.py 262144 2>&1 | tail -1
done
echo "== generic"; NGRTD_NG_GENERIC=1 NGRTD_LIB=$PWD/build_exp/lib_ngr7.so python tools/ng_probe2.py 2>&1 | tail -1
NGRTD_NG_GENERIC=1 NGRTD_LIB=$PWD/build_exp/lib_ngr7.so python tools/ng_probe2.py 262144 2>&1 | tail -1
NGRTD_LIB=$PWD/build_exp/lib_ngr7.so timeout 600 python -m pytest tests/test_sampler_gpu.py tests/test_r2_gpu.py -m gpu -q -k "ng_ or pooled or shard_inv or checkpoint or observation_groups" 2>&1 | tail -5
v=ngr7
NGRTD_LIB=$PWD/build_exp/lib_$v.so timeout 300 ncu --set full --import-source on --clock-control none -k regex:k_mcmc_ng -s 2 -c 1 -o gpurun_out/r2g_$v python tools/prof_ng.py > gpurun_out/r2g_ncu_$v.log 2>&1
ncu -i gpurun_out/r2g_$v.ncu-rep --page raw --csv > gpurun_out/r2g_${v}_raw.csv 2>/dev/null
ncu -i gpurun_out/r2g_$v.ncu-rep --page source --csv > gpurun_out/r2g_${v}_src.csv 2>/dev/null
python tools/ncu_summary.py gpurun_out/r2g_${v}_raw.csv gpurun_out/r2g_${v}_src.csv > gpurun_out/r2g_${v}_summary.txt 2>/dev/null; grep -E "time_duration|long_scoreboard|fp64_cycles|dram__bytes|registers|warps_active" gpurun_out/r2g_${v}_summary.txt;  grep -A10 "top instructions" gpurun_out/r2g_${v}_summary.txt
rm -f gpurun_out/r2g_$v.ncu-rep
