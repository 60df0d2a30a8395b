#!/bin/bash
# sampler kernel with the 2,048-entry exp table (needs the per-chain records shrunk to fit shared memory)
for l in s7 s7_nd8 s11 s7 s11; do NGRTD_LIB=$PWD/build_exp/$l.so python tools/sampler_time.py 2>&1 | tail -1; done
