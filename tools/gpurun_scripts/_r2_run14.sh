#!/bin/bash
timeout 900 python examples/config4_joint_fit.py 49152 256 10000 100000 2>&1 | grep -E "R-hat|99th|groups with|sampling"
timeout 900 python examples/config4_joint_fit.py 49152 256 20000 200000 2>&1 | grep -E "R-hat|99th|groups with|sampling"
