#!/bin/bash
# more seeds of the randomised parity sweep (final library)
for s in 101 102 103 104 105 106; do timeout 600 python tools/fuzz_forward.py 500 $s 2>&1 | tail -4; done
