#!/bin/bash
for l in noblegas_rtd_mcmc_b200/libngrtd.so build_exp/lib_split.so; do NGRTD_LIB=$PWD/$l python tools/r2_probe.py 65536,131072 2>&1 | tail -1; NGRTD_LIB=$PWD/$l python tools/sampler_time.py 2>&1 | tail -1; done
