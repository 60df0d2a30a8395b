#!/bin/bash
mkdir -p gpurun_out
for lib in lib_r1 lib_exp; do
  NGRTD_LIB=$PWD/build_exp/$lib.so timeout 120 python tools/variant_bench.py 2>&1 | tail -3
done > gpurun_out/r2_variant1.txt 2>&1
NGRTD_PDL=0 NGRTD_LIB=$PWD/build_exp/lib_exp.so timeout 120 python tools/variant_bench.py 2>&1 | tail -3 >> gpurun_out/r2_variant1.txt
cat gpurun_out/r2_variant1.txt
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 > gpurun_out/r2_pytest1.txt
cat gpurun_out/r2_pytest1.txt
