#!/bin/bash
mkdir -p gpurun_out
L=$PWD/build_exp
for lib in lib_exp lib_w20 lib_w24 lib_rep2 lib_w20rep2; do
NGRTD_LIB=$L/$lib.so timeout 120 python tools/r2_probe.py 65536,131072,303104
NGRTD_LIB=$L/$lib.so timeout 120 python tools/variant_bench.py 2>&1 | tail -1
done > gpurun_out/r2_probe3.txt 2>&1
cat gpurun_out/r2_probe3.txt
