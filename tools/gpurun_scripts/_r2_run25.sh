#!/bin/bash
for v in ngr8 ngr8g; do for c in -2 -1 25 50 100; do
  echo "== $v carve $c"; NGRTD_NG_CARVE=$c NGRTD_LIB=$PWD/build_exp/lib_$v.so python tools/ng_probe2.py 2>&1 | tail -1
done; done
