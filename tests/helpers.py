"""Shared fixtures for the parity tests (golden loading, tracer tables)."""
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
J_MU = -10.424151075511698

MODEL_CFGS = {
    "emm0": ("exponential", False, ["tau1"]),
    "emm123": ("exponential", False, ["tau1", "J", "thalf_cfc", "lamsf6"]),
    "epm123": ("exp_pist_flow", False, ["tau1", "eta1", "J", "thalf_cfc", "lamsf6"]),
    "epm_pfm123": ("exp_pist_flow", "piston", ["tau1", "tau2", "f1", "f2", "eta1", "J", "thalf_cfc", "lamsf6"]),
    "emm_pfm123": ("exponential", "piston", ["tau1", "tau2", "f1", "f2", "J", "thalf_cfc", "lamsf6"]),
    "dm": ("dispersion", False, ["tau1", "D1"]),
    "pfm": ("piston", False, ["tau1"]),
    "epm_dm": ("exp_pist_flow", "dispersion", ["tau1", "tau2", "f1", "f2", "eta1", "D2", "J"]),
    "dm_emm": ("dispersion", "exponential", ["tau1", "tau2", "f1", "f2", "D1", "J", "lamsf6"]),
    "dm_dm": ("dispersion", "dispersion", ["tau1", "tau2", "f1", "f2", "D1", "D2"]),
    "pfm_epm": ("piston", "exp_pist_flow", ["tau1", "tau2", "f1", "f2", "eta2", "thalf_cfc"]),
}

from noblegas_rtd_mcmc_b200.datasets import REAL_TRACERS, load_c_in  # noqa: E402,F401  (the series live in the package)


def rel_err(a, b):
    """Max relative error with NaN/inf patterns required to agree."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    na, nb = np.isnan(a), np.isnan(b)
    assert np.array_equal(na, nb), "NaN pattern differs: %d vs %d" % (na.sum(), nb.sum())
    m = ~na
    if not m.any():
        return 0.0
    inf = np.isinf(a[m]) | np.isinf(b[m])
    assert np.array_equal(a[m][inf], b[m][inf])
    aa, bb = a[m][~inf], b[m][~inf]
    den = np.maximum(np.abs(bb), 1e-300)
    return float(np.max(np.abs(aa - bb) / den)) if aa.size else 0.0


def synth_descs(par_names, tracers=None, L=840, seed=0):
    from noblegas_rtd_mcmc_b200 import synthetic
    return synthetic.series_matrix_and_descs(par_names, tracers, L, seed)


def synth_plan(mod1, mod2, par_names, tracers=None, L=840, seed=0, device=-1):
    """Plan over the synthetic monthly series for the given tracers (default: the seven cfg-3 tracers)."""
    from noblegas_rtd_mcmc_b200 import _lib, synthetic
    X, descs = synth_descs(par_names, tracers, L, seed)
    return _lib.Plan(X, descs, mod1, mod2, device=device), synthetic.input_series(L, seed), synthetic.tracer_table_cfg3()


def real_plan(mod1, mod2, par_names, tracers, L=None):
    """Plan over the reference's yearly series (rebuilt from the packaged head + constant background)."""
    from noblegas_rtd_mcmc_b200 import _lib, datasets
    X, descs, C = datasets.real_series_matrix_and_descs(par_names, tracers, L)
    return _lib.Plan(X, descs, mod1, mod2), C
