#!/bin/bash
NGRTD_LIB=$PWD/build_exp/s_fin.so python tools/sampler_time.py 2>&1 | tail -1
NGRTD_LIB=$PWD/build_exp/s_fin.so NGRTD_MCMC_TB11=0 python tools/sampler_time.py 2>&1 | tail -1
