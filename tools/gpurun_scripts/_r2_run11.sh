#!/bin/bash
mkdir -p gpurun_out
python tools/r2_debug_tape.py 2>&1 | grep -E "^B|bad"
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15 > gpurun_out/r2_pytest11.txt
cat gpurun_out/r2_pytest11.txt
NGRTD_LIB=$PWD/build_exp/lib_exp.so timeout 120 python tools/variant_bench.py 2>&1 | tail -1
