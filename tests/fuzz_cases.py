"""Randomised parity cases of the forward model against the numpy oracle (reference arithmetic): random lag-axis lengths,
batch sizes, model pairs, tracer sets (decay, 3He ingrowth, 4He accumulation, per-chain CFC decay, SF6 excess), series with
and without a constant tail, wide parameter ranges.  Used by tests/test_fuzz_gpu.py and tools/fuzz_forward.py (GPU).
Chains whose largest dispersion exponent lies in the documented deviation band (e^-800 .. e^-650 of the window maximum:
the reference returns numbers built from denormals there, the kernel NaN) are excluded from the comparison."""
import numpy as np

import np_oracle as O
from noblegas_rtd_mcmc_b200 import _lib

MODS = ["piston", "exponential", "exp_pist_flow", "dispersion"]


def emax_dispersion(tau, D, L):
    """largest exponent -(tp - tau)^2 / (4 D tau tp) over the lag grid"""
    tp = O.lag_grid(L)
    k = np.clip(np.floor(tau).astype(int), 0, L - 1)
    k2 = np.clip(k + 1, 0, L - 1)
    e = lambda t: -(t - tau) ** 2 / (4 * D * tau * t)
    return np.maximum(e(tp[k]), e(tp[k2]))


def one_case(rng, case):
    L = int(rng.choice([1, 2, 3, 4, 5, 7, 8, 31, 64, 100, 257, 840, 1023, 1024, 1025, 2000, 3001]))
    B = int(rng.choice([1, 2, 7, 8, 9, 15, 16, 17, 33, 100, 511, 1000, 2368 * 16 // 16, 4099, 6000]))
    m1 = MODS[rng.integers(0, 4)]
    m2 = [False] + MODS
    m2 = m2[rng.integers(0, 5)]
    nser = 3
    X = rng.uniform(0.5, 50.0, (L, nser))
    tail = L >= 64 and rng.random() < 0.5
    if tail:                                           # constant background beyond a random cut (the reference's back-extension)
        cut = int(rng.integers(4, max(5, L // 2)))
        X[cut:] = rng.uniform(0.1, 5.0, nser)
    dyn = rng.random() < 0.4
    lam = float(np.log(2) / rng.uniform(5, 40))
    descs = [dict(series=0, lam=lam), dict(series=0, lam=lam, rad_accum="3He"), dict(series=1),
             dict(series=-1, rad_accum="4He"), dict(series=1, use_lamsf6=True)]
    names = ["H3", "He3", "X", "He4_ter", "SF6"]
    if dyn:
        descs.append(dict(series=2, use_thalf_cfc=True)); names.append("CFC12")
    ntr = int(rng.integers(1, len(descs) + 1))
    sel = sorted(rng.choice(len(descs), ntr, replace=False).tolist())
    descs = [descs[i] for i in sel]; names = [names[i] for i in sel]
    pn = ["tau1", "J", "lamsf6", "thalf_cfc"]
    if m1 in ("exp_pist_flow",): pn.append("eta1")
    if m1 == "dispersion": pn.append("D1")
    if m2:
        pn += ["tau2", "f1", "f2"]
        if m2 == "exp_pist_flow": pn.append("eta2")
        if m2 == "dispersion": pn.append("D2")
    f1 = rng.uniform(0.02, 0.98, B)
    span = max(2.0, float(L))
    cols = {"tau1": np.exp(rng.uniform(np.log(0.3), np.log(30 * span), B)), "tau2": np.exp(rng.uniform(np.log(0.3), np.log(30 * span), B)),
            "f1": f1, "f2": 1 - f1, "eta1": rng.uniform(1, 5, B), "eta2": rng.uniform(1, 5, B),
            "D1": np.exp(rng.uniform(np.log(0.01), np.log(3.0), B)), "D2": np.exp(rng.uniform(np.log(0.01), np.log(3.0), B)),
            "J": rng.normal(-10.4, 0.4, B), "lamsf6": np.abs(rng.normal(0, 0.17, B)), "thalf_cfc": rng.uniform(5, 60, B)}
    theta = np.stack([cols[p] for p in pn], axis=1)
    keep = np.ones(B, bool)
    for m, t, d in ((m1, "tau1", "D1"), (m2, "tau2", "D2")):
        if m == "dispersion":
            em = emax_dispersion(cols[t], cols[d], L)
            keep &= ~((em < -650.0) & (em > -800.0))
    plan = _lib.Plan(X, descs, m1, m2)
    got = plan.forward_host(theta, pn)
    want = np.empty_like(got)
    for i, d in enumerate(descs):
        s = X[:, d["series"]] if d.get("series", -1) >= 0 else np.zeros(L)
        with np.errstate(all="ignore"):
            want[:, i] = O.forward_mod(theta, pn, names[i], s, m1, m2, t_half=(np.log(2) / d["lam"] if d.get("lam") else False),
                                       rad_accum=d.get("rad_accum", False))
    g, w = got[keep], want[keep]
    nan_g, nan_w = np.isnan(g), np.isnan(w)
    bad_nan = int((nan_g != nan_w).sum())
    fin = ~(nan_g | nan_w) & np.isfinite(w) & np.isfinite(g)
    # 3He ingrowth at lag 0 cancels 7 digits in the reference itself (1 - exp(-lambda 1e-5)): looser bound where it dominates
    # (any lag axis when a piston component sits at lag 0, tau < 0.5; short axes for every model)
    lag0 = np.zeros(B, bool)
    for m, t in ((m1, "tau1"), (m2, "tau2")):
        if m == "piston":
            lag0 |= cols[t] < 0.5
    he3 = np.array([n == "He3" for n in names])[None, :]
    tol = np.where(he3 & ((L <= 4) | lag0[keep][:, None]), 1e-7, 1e-10) * np.ones_like(w)
    err = np.abs(g - w) / np.maximum(np.abs(w), 1e-300)
    worst = float(np.max(np.where(fin, err / tol, 0.0))) if fin.any() else 0.0
    tag = "L=%d B=%d %s+%s tail=%d%s dyn=%d tracers=%s" % (L, B, m1, m2, tail, ("(cut %d)" % cut) if tail else "", dyn, ",".join(names))
    if fin.any() and float(np.max(np.where(fin, err / tol, 0.0))) > 1.0:      # the worst chains, for the report
        rows = np.argsort(-np.max(np.where(fin, err / tol, 0.0), axis=1))[:3]
        kept = np.flatnonzero(keep)
        for r in rows:
            tag += "\n      chain %d: %s  err %s" % (kept[r], " ".join("%s=%.6g" % (p, theta[kept[r], j]) for j, p in enumerate(pn)),
                                                    np.array2string(np.where(fin[r], err[r], np.nan), precision=2))
    return bad_nan, worst, tag, int(keep.sum()), float(np.max(np.where(fin, err, 0.0))) if fin.any() else 0.0
