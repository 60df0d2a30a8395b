#!/bin/bash
# sampler kernel: 2,048-entry exp table when it fits (per-chain records sized by nd) -- timing A/B by env switch, sampler tests
python tools/sampler_time.py 2>&1 | tail -1
NGRTD_MCMC_TB11=0 python tools/sampler_time.py 2>&1 | tail -1
timeout 1500 python -m pytest tests -m gpu -q -x -k "sampler or bmm or mcmc or posterior or trace or population or config" 2>&1 | tail -3
