#!/bin/bash
# A/B: r from bits (one DFMA less per dispersion weight) against the current build
for l in v0 v_r32 v0 v_r32; do NGRTD_LIB=$PWD/build_exp/$l.so timeout 300 python tools/variant_bench.py 2>&1 | tail -1; done
