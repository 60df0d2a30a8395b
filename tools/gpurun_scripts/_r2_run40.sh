#!/bin/bash
NGRTD_LIB=$PWD/build_exp/v_tune.so timeout 600 python tools/tune_sweep2.py 2>&1 | tail -12
