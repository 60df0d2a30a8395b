#!/bin/bash
# does nvcc -split-compile change the kernels? (build time 4.4 -> 2.1 min)
NGRTD_LIB=$PWD/build_exp/full_split.so python tools/variant_bench.py 2>&1 | tail -1
python tools/variant_bench.py 2>&1 | tail -1
NGRTD_LIB=$PWD/build_exp/full_split.so python tools/sampler_time.py 2>&1 | tail -1
python tools/sampler_time.py 2>&1 | tail -1
