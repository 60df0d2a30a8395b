"""Randomised parity sweep on the GPU: the forward model through the C ABI against the numpy oracle over random lag-axis
lengths (1 .. 3001), batch sizes (1 .. 6000: every tape cut pattern), model pairs, tracer sets and parameter ranges
(tests/fuzz_cases.py).  Bound: 1e-10 relative, identical NaN patterns (1e-7 for 3He where the reference itself cancels
7 digits at lag 0)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed", [1, 2, 7])
def test_forward_fuzz_against_oracle(seed):
    from fuzz_cases import one_case
    rng = np.random.default_rng(seed)
    fails = []
    for c in range(120):
        bad_nan, worst, tag, nkeep, e = one_case(rng, c)
        if bad_nan or worst > 1.0:
            fails.append("case %d: %s | NaN mismatches %d, worst err/tol %.3g" % (c, tag, bad_nan, worst))
    assert not fails, "\n".join(fails)
