#!/bin/bash
python tools/bmm_posterior_probe.py 128 2>&1 | grep -E "pool|IS|tau1|eta1" 
timeout 600 python examples/config4_joint_fit.py 49152 256 10000 10000 2>&1 | grep -E "R-hat|99th|groups with|sampling"
