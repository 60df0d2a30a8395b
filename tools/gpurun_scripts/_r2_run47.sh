#!/bin/bash
for l in v_skip2 v_skip3; do NGRTD_LIB=$PWD/build_exp/$l.so python tools/dm_tail_time.py 2>&1 | tail -1; done
