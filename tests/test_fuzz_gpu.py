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


@pytest.mark.parametrize("S", [0.0, 5.0, 35.0])
def test_ce_fuzz_against_oracle(S):
    """Closed-equilibrium model (k_ce) over 20,000 random parameter sets per salinity against the numpy oracle: recharge
    elevations 0 .. 5,000 m, temperatures -2 .. 110 C (the T < 0 sentinel and both Antoine branches), Ae over 6 decades,
    F 0 .. 2; bound 1e-12 relative (times the cancellation factor of P - P_v), identical NaN / sentinel patterns."""
    import np_oracle as O
    from noblegas_rtd_mcmc_b200.noble_gas_utils import noble_gas_fun
    rng = np.random.default_rng(int(S) + 3)
    n = 20000
    gases = ["He", "Ne", "Ar", "Kr", "Xe"]
    E = rng.uniform(0.0, 5000.0, n)
    T = rng.uniform(-2.0, 110.0, n)
    Ae = 10.0 ** rng.uniform(-6.0, 0.0, n)
    F = rng.uniform(0.0, 2.0, n)
    # C_eq is proportional to P - P_v, which cancels where the water would boil (hot recharge at altitude: P_v ~ P): there the
    # last-ulp differences between CUDA's and numpy's pow / exp10 are amplified by P / |P - P_v| in the reference's own formula
    P, Pv = O.lapse_rate(E), O.vapor_pressure(T)
    with np.errstate(all="ignore"):
        amp = np.where(T >= 0.0, np.maximum(1.0, P / np.abs(P - Pv)), 1.0)
    for add in (True, False):
        got = noble_gas_fun(gases, E, T, Ae, F, "lapse_rate", S=S).ce_exc(add)
        want = O.ce_exc(gases, E, T, Ae, F, add, S=S)
        for i, g in enumerate(gases):
            assert np.array_equal(np.isnan(got[g]), np.isnan(want[:, i])), g
            with np.errstate(all="ignore"):
                err = np.abs(got[g] - want[:, i]) / np.abs(want[:, i])
            err = np.where(np.isfinite(err), err, 0.0)
            assert np.all(err <= 1e-12 * amp), (g, add, float(np.max(err / amp)))
