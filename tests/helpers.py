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

# tracer -> (series key, t_half, rad_accum) on the real yearly data (reference run_age_mcmc.py:200-224)
REAL_TRACERS = {
    "CFC11": ("CFC11", False, False),
    "CFC12": ("CFC12", False, False),
    "CFC113": ("CFC113", False, False),
    "SF6": ("SF6", False, False),
    "He4_ter": (None, False, "4He"),
    "He3": ("H3", 12.34, "3He"),
    "H3": ("H3", 12.34, False),
}


def load_c_in(L=None):
    """Rebuild the reference's C_in_dict series (newest-first float64[L]) from the committed head."""
    z = np.load(os.path.join(GOLD, "c_in_head.npz"))
    Lfull = int(z["L"])
    L = Lfull if L is None else L
    out = {}
    for k in ("CFC11", "CFC12", "CFC113", "SF6", "He4_ter", "H3"):
        v = np.full(L, float(z[k + "_bg"]))
        n = min(L, 128)
        v[:n] = z[k + "_head"][:n]
        out[k] = v
    return out


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
    """Plan over the reference's yearly series (rebuilt from the committed head + constant background)."""
    from noblegas_rtd_mcmc_b200 import _lib
    C = load_c_in(L)
    names = ["CFC11", "CFC12", "CFC113", "SF6", "H3"]
    X = np.stack([C[n] for n in names], axis=1)
    descs = []
    for t in tracers:
        s, th, ra = REAL_TRACERS[t]
        descs.append(dict(series=names.index(s) if s is not None else -1, rad_accum=ra,
                          lam=float(-1.0 * np.log(0.5) / th) if th else 0.0,
                          use_thalf_cfc=(t == "CFC12" and "thalf_cfc" in par_names),
                          use_lamsf6=(t == "SF6")))
    return _lib.Plan(X, descs, mod1, mod2), C
