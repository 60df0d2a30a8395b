#!/bin/bash
# A/B of build variants (session 3): exp table 7 vs 11 bits, tp carried by DADD, swizzled Xf
for l in v0 v_tb7 v_dadd v_swz v_swz_dadd v0; do NGRTD_LIB=$PWD/build_exp/$l.so timeout 300 python tools/variant_bench.py 2>&1 | tail -1; done
