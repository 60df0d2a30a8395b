#!/bin/bash
# schedule perturbations of the (G, D) lag loop + the randomised parity test
for l in v0 p1 p2 p4 p6 v0; do NGRTD_LIB=$PWD/build_exp/$l.so timeout 300 python tools/variant_bench.py 2>&1 | tail -1; done
timeout 900 python -m pytest tests/test_fuzz_gpu.py -q -x -m gpu 2>&1 | tail -5
