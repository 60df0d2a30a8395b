#!/bin/bash
mkdir -p gpurun_out
L=$PWD/build_exp
for lib in "$@"; do
NGRTD_LIB=$L/$lib.so timeout 120 python tools/variant_bench.py 2>&1 | tail -1
done | tee gpurun_out/r2_variant5.txt
