#!/bin/bash
for l in lib_base.so lib_exp.so lib_base.so lib_exp.so; do NGRTD_LIB=$PWD/build_exp/$l python tools/r2_probe.py 65536,131072 2>&1 | tail -1; done
NGRTD_LIB=$PWD/build_exp/lib_exp.so timeout 900 python -m pytest tests -m gpu -q -x -k "golden or forward or dropin or fullsize or tape or tail or dm_ or extreme" 2>&1 | tail -4
