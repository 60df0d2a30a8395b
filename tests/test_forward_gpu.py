"""GPU parity: the CUDA forward path (through the C ABI) against the golden vectors of the untouched
reference and against the numpy oracle on seeded draws.  Gate: 1e-10 relative (north_star), NaN patterns equal."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

TOL = 1e-10


def _lib():
    from noblegas_rtd_mcmc_b200 import _lib
    return _lib


@pytest.mark.parametrize("name", ["epm_dm", "emm0", "dm", "dm_dm", "epm_pfm123", "pfm", "cfg3", "cfg3i"])
def test_forward_synthetic_vs_reference_golden(name):
    from helpers import GOLD, MODEL_CFGS, rel_err, synth_plan
    from noblegas_rtd_mcmc_b200 import synthetic
    z = np.load(os.path.join(GOLD, "forward_synth.npz"))
    if name.startswith("cfg3"):
        m1, m2, pn = "exp_pist_flow", "dispersion", list(synthetic.PAR_NAMES_CFG3)
    else:
        m1, m2, pn = MODEL_CFGS[name]
    plan, _, _ = synth_plan(m1, m2, pn)
    theta = z[name + "/theta"]
    out = plan.forward_host(theta, pn)
    for i, t in enumerate(synthetic.TRACERS_CFG3):
        e = rel_err(out[:, i], z[name + "/" + t])
        assert e < TOL, (name, t, e)


@pytest.mark.parametrize("name", sorted(__import__("helpers").MODEL_CFGS))
def test_forward_real_series_vs_reference_golden(name):
    """L = 25,256 yearly lags: exercises the chunk-streamed (lock-step) path."""
    from helpers import GOLD, MODEL_CFGS, rel_err, real_plan
    z = np.load(os.path.join(GOLD, "forward_real.npz"))
    m1, m2, pn = MODEL_CFGS[name]
    tracers = ["CFC12", "SF6", "H3", "He4_ter", "He3", "CFC11"]
    plan, _ = real_plan(m1, m2, pn, tracers)
    theta = z[name + "/theta"]
    out = plan.forward_host(theta, pn)
    for i, t in enumerate(tracers):
        e = rel_err(out[:, i], z[name + "/" + t])
        assert e < TOL, (name, t, e)


def test_forward_vs_oracle_seeded_batch():
    """4,099 chains (ragged tile tail) of the cfg-3 mixture against the numpy oracle."""
    import np_oracle as O
    from helpers import rel_err, synth_plan
    from noblegas_rtd_mcmc_b200 import synthetic
    pn = list(synthetic.PAR_NAMES_CFG3)
    plan, series, tab = synth_plan("exp_pist_flow", "dispersion", pn)
    theta = synthetic.theta_cfg3_informative(4099, 5)
    out = plan.forward_host(theta, pn)
    worst = 0.0
    for i, t in enumerate(synthetic.TRACERS_CFG3):
        d = tab[t]
        s = series[d["series"]] if d["series"] is not None else np.zeros(840)
        want = O.forward_mod(theta, pn, t, s, "exp_pist_flow", "dispersion", t_half=d.get("t_half", False),
                             rad_accum=d.get("rad_accum", False))
        worst = max(worst, rel_err(out[:, i], want))
    assert worst < TOL, worst
