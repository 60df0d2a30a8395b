#!/bin/bash
for l in v0 v_setup v0 v_setup; do NGRTD_LIB=$PWD/build_exp/$l.so timeout 300 python tools/variant_bench.py 2>&1 | tail -1; done
