#!/bin/bash
# 12 warps (168 registers) with deeper unrolling of the lag loop
for l in t_u3 t_u4; do echo $l; NGRTD_LIB=$PWD/build_exp/$l.so timeout 600 python tools/tune_sweep2.py 2>&1 | grep -E "warps=0|warps=12|warps=8  NT=4"; done
