#!/bin/bash
for l in noblegas_rtd_mcmc_b200/libngrtd.so build_exp/lib_st.so build_exp/lib_r1.so; do NGRTD_LIB=$PWD/$l python tools/sampler_time.py 2>&1 | tail -1; done
