"""Shared set-up of the two-component (binary mixing model) posterior checks: the reference's `.123` exp_pist_flow-piston and
exponential-piston inversions of well PLM1 (priors run_age_mcmc_utils.py:286-344, bounds run_age_mcmc.py:144-178), observations
from the reference's own trace file, errors from tests/golden/age_obs_err.json.

The reference's traces of these models are not converged (R-hat 1.25-1.56, bulk-ESS 5-9), so the yardstick is the EXACT
posterior: self-normalised importance sampling from the prior with ~1e9 draws evaluated by the (golden-pinned) forward +
Student-T likelihood kernel -- exact up to its own Monte-Carlo error, which is estimated and used in the bounds."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
TRACERS = ["CFC12", "SF6", "H3", "He4_ter"]


def setup(model1, well="PLM1"):
    from helpers import real_plan
    from noblegas_rtd_mcmc_b200 import noble_gas_utils as ng_utils
    from noblegas_rtd_mcmc_b200.sampler import prior
    fx = json.load(open(os.path.join(ROOT, "tests", "golden", "age_traces.json")))
    rel = json.load(open(os.path.join(ROOT, "tests", "golden", "age_obs_err.json")))["rel"]
    key = "%s.CFC12.SF6.H3.He4_ter.exp_pist_flow.123" % well          # same observations in every joint trace of a well
    obs = np.array(fx["traces"][key]["obs_mu"])
    sd = np.array([rel[t][well] for t in TRACERS]) * obs
    epm = model1 == "exp_pist_flow"
    pn = ["tau1", "tau2", "f1", "f2"] + (["eta1"] if epm else []) + ["J", "thalf_cfc", "lamsf6"]
    plan, _ = real_plan(model1, "piston", pn, TRACERS)
    J_mu = np.log10(ng_utils.J_flux(Del=1., rho_r=2700, rho_w=1000, U=3.7, Th=10.2, phi=0.05))
    pri = [prior("uniform", "tau1", 1.0, 1000.0), prior("beta", "nu_", 2.0, 0.1), prior("normal", "J", J_mu, 0.33),
           prior("uniform", "tau2", 50.0, 15000.0), prior("uniform", "f1", 0.01, 0.99)]
    if epm:
        pri.append(prior("uniform", "eta1", 1.0, 5.0))
    pri += [prior("beta", "thalf_cfc", 2.0, 2.0, lo=5.0, hi=35.0), prior("halfnormal", "lamsf6", 0.5 / 3)]
    return plan, pn, pri, obs, sd, J_mu


def exact_posterior(plan, pn, obs, sd, J_mu, n_batches=1024, batch=1 << 20, seed=0, keep_below_max=30.0):
    """Importance sampling from the prior on the device.  Returns (theta_kept [n, ndim+1] natural values incl. nu_ as the
    last column, normalised weights, ESS)."""
    import torch
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    dev = torch.device("cuda")
    epm = "eta1" in pn
    beta22 = torch.distributions.Beta(torch.tensor(2.0, device=dev, dtype=torch.float64), torch.tensor(2.0, device=dev, dtype=torch.float64))
    beta_nu = torch.distributions.Beta(torch.tensor(2.0, device=dev, dtype=torch.float64), torch.tensor(0.1, device=dev, dtype=torch.float64))
    torch.manual_seed(seed)
    kept_t, kept_w = [], []
    gmax = -np.inf
    for b in range(n_batches):
        u = torch.rand((batch, 4), generator=g, device=dev, dtype=torch.float64)
        z = torch.randn((batch, 2), generator=g, device=dev, dtype=torch.float64)
        tau1 = 1.0 + 999.0 * u[:, 0]
        tau2 = 50.0 + 14950.0 * u[:, 1]
        f1 = 0.01 + 0.98 * u[:, 2]
        cols = {"tau1": tau1, "tau2": tau2, "f1": f1, "f2": 1.0 - f1, "eta1": 1.0 + 4.0 * u[:, 3], "J": J_mu + 0.33 * z[:, 0],
                "thalf_cfc": 5.0 + 30.0 * beta22.sample((batch,)), "lamsf6": (0.5 / 3) * z[:, 1].abs()}
        nu_ = beta_nu.sample((batch,))
        theta = torch.stack([cols[p] for p in pn], dim=1).contiguous()
        nu = (5.0 + 25.0 * nu_).contiguous()
        lw = plan.forward_loglik_dev(theta, pn, obs, sd, "studentt", nu_t=nu)
        lw = torch.where(torch.isfinite(lw), lw, torch.full_like(lw, -1e300))
        m = float(lw.max())
        gmax = max(gmax, m)
        sel = lw > gmax - keep_below_max
        if bool(sel.any()):
            kept_t.append(torch.cat([theta[sel], nu_[sel, None]], dim=1).cpu().numpy())
            kept_w.append(lw[sel].cpu().numpy())
    th = np.concatenate(kept_t)
    lw = np.concatenate(kept_w)
    ok = lw > gmax - keep_below_max
    th, lw = th[ok], lw[ok]
    w = np.exp(lw - lw.max())
    w /= w.sum()
    return th, w, 1.0 / np.sum(w * w)


def weighted_cdf(x, w, q):
    o = np.argsort(x)
    cw = np.cumsum(w[o])
    return np.interp(q, x[o], cw)
