#!/bin/bash
python tools/r2_debug_tape.py 2>&1 | tail -40
echo "== no split"; NGRTD_TAPE_MIN=100000 python tools/r2_debug_tape.py 2>&1 | grep -E "^B|bad" 
echo "== stage1"; NGRTD_STAGE=1 python tools/r2_debug_tape.py 2>&1 | grep -E "^B|bad"
