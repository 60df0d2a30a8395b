#!/bin/bash
# the examples on the final library (the age sampler now runs the 2,048-entry-table instantiation where it fits)
for e in config1_noble_gas_fit config2_age_fit; do echo "== $e"; timeout 600 python examples/$e.py 2>&1 | tail -12; done
echo "== config4_joint_fit (1 GPU)"; timeout 900 python examples/config4_joint_fit.py 2>&1 | tail -8
cc -O2 -Iinclude examples/c_abi_demo.c -o /tmp/c_abi_demo -Lnoblegas_rtd_mcmc_b200 -lngrtd -Wl,-rpath,$PWD/noblegas_rtd_mcmc_b200 -lm 2>&1 | tail -2 && /tmp/c_abi_demo | tail -3
