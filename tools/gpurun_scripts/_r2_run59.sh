#!/bin/bash
# dispersion tail: packed column types, 1/t from rsqrt -- timing A/B and parity on the real-series goldens
for l in d0 d1 d0 d1; do NGRTD_LIB=$PWD/build_exp/$l.so python tools/dm_tail_time.py 2>&1 | tail -1; done
NGRTD_LIB=$PWD/build_exp/d1.so python tools/dm_tail_check.py 2>&1 | tail -8
