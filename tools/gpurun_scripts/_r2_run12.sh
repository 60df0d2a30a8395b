#!/bin/bash
mkdir -p gpurun_out
nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -shared -Xcompiler -fPIC -o noblegas_rtd_mcmc_b200/libngrtd.so noblegas_rtd_mcmc_b200/csrc/ngrtd_api.cu
timeout 900 python -m pytest tests/test_bmm_posterior_gpu.py tests/test_sampler_gpu.py -q 2>&1 | tail -15
timeout 600 python examples/config4_joint_fit.py 49152 256 10000 10000 2>&1 | tail -30 | tee gpurun_out/r2_config4_1gpu.txt
