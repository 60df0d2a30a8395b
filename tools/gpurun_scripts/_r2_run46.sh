#!/bin/bash
for l in v_head v_skip v_skip2; do NGRTD_LIB=$PWD/build_exp/$l.so python tools/dm_tail_time.py 2>&1 | tail -1; done
NGRTD_LIB=$PWD/build_exp/v_skip2.so python tools/gpurun_scripts/_dbg_extreme.py 2>&1 | tail -12
