"""noblegas_rtd_mcmc_b200 -- B200-native (sm_100a) likelihood hot path of uz226/NobleGas_RTD_MCMC.

Importing the package never touches the GPU; the CUDA library (libngrtd.so, C ABI in include/ngrtd.h) is loaded
by `noblegas_rtd_mcmc_b200._lib` on first use of an operator and there is no CPU fallback.
"""
__version__ = "0.1.0"
