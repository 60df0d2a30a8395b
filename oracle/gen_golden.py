"""TEST INFRASTRUCTURE ONLY -- generate golden vectors by running the UNTOUCHED reference
(/root/reference) in the build container through oracle/ref_shims.py.

    python oracle/gen_golden.py          # writes tests/golden/*.npz

The reference cannot travel to the GPU box, so the vectors are committed as small fixtures.
Inputs that are data of the reference (C_in_dict.pk: 25,256-lag series whose oldest >= 25,171
rows are one constant) are stored as `head` (newest 128 lags) + background.
"""
import os
import sys

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_shims  # noqa: E402
from noblegas_rtd_mcmc_b200 import synthetic  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
J_MU = -10.424151075511698

# model configurations: (mod_type1, mod_type2, par_names)  -- SURVEY App. B table + DM variants
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


def draw_theta(par_names, n, rng, tau_scale=1.0):
    cols = []
    f1 = rng.uniform(0.01, 0.99, n)
    for p in par_names:
        if p == "tau1":
            v = np.exp(rng.uniform(np.log(1.0), np.log(1000.0), n)) * tau_scale
        elif p == "tau2":
            v = np.exp(rng.uniform(np.log(50.0), np.log(15000.0), n)) * tau_scale
        elif p == "f1":
            v = f1
        elif p == "f2":
            v = 1.0 - f1
        elif p in ("eta1", "eta2"):
            v = rng.uniform(1.0, 5.0, n)
        elif p in ("D1", "D2"):
            v = rng.uniform(0.01, 2.0, n)
        elif p == "J":
            v = rng.normal(J_MU - np.log10(tau_scale), 0.33, n)
        elif p == "thalf_cfc":
            v = rng.uniform(5.0, 35.0, n) * tau_scale
        elif p == "lamsf6":
            v = np.abs(rng.normal(0.0, 0.5 / 3, n))
        else:
            raise KeyError(p)
        cols.append(v)
    return np.stack(cols, axis=1)


def ckw_real(C_in, tracer):
    """run_age_mcmc.py:200-224 of the reference."""
    kw = {}
    if tracer in ("CFC11", "CFC12", "CFC113", "SF6"):
        kw["C_t"] = C_in[tracer].copy()
    elif tracer == "He4_ter":
        kw["C_t"] = C_in[tracer].copy() * 0.0
        kw["rad_accum"] = "4He"
    elif tracer == "He3":
        kw["C_t"] = C_in["H3"].copy()
        kw["t_half"] = 12.34
        kw["rad_accum"] = "3He"
    elif tracer == "H3":
        kw["C_t"] = C_in["H3"].copy()
        kw["t_half"] = 12.34
    return kw


def ckw_synth(series, tracer):
    L = len(series["H3"])
    idx = np.arange(L - 1, -1, -1)
    tab = synthetic.tracer_table_cfg3()[tracer]
    kw = {}
    if tab["series"] is None:
        kw["C_t"] = pd.DataFrame({tracer: np.zeros(L)}, index=idx)
    else:
        kw["C_t"] = pd.DataFrame({tracer: series[tab["series"]][::-1].copy()}, index=idx)
    if "t_half" in tab:
        kw["t_half"] = tab["t_half"]
    if "rad_accum" in tab:
        kw["rad_accum"] = tab["rad_accum"]
    return kw


def run_forward(ramu, kw, mod1, mod2, par_names, tracer, theta):
    kw = dict(kw)
    kw["mod_type1"], kw["mod_type2"] = mod1, mod2
    op = ramu.ForwardMod(kw, par_names, tracer)
    out = np.empty(theta.shape[0])
    with np.errstate(all="ignore"):
        for i in range(theta.shape[0]):
            o = [[None]]
            op.p_dict = dict(op.p_dict)
            op.perform(None, [theta[i]], o)
            out[i] = float(o[0][0])
    return out


def main():
    os.makedirs(GOLD, exist_ok=True)
    conv, ngu, ramu = ref_shims.load()
    C_in = pd.read_pickle(os.path.join(ref_shims.REFERENCE_ROOT, "age_ens_runs_mcmc", "C_in_dict.pk"))
    rng = np.random.default_rng(20261018)

    # ---- input-series fixture (newest-first head + constant background) ----
    head = {}
    for k, df in C_in.items():
        v = df.to_numpy().ravel()[::-1]
        assert np.all(v[128:] == v[128]), k
        head[k + "_head"] = v[:128].copy()
        head[k + "_bg"] = np.array(v[128])
    head["L"] = np.array(len(v))
    np.savez_compressed(os.path.join(GOLD, "c_in_head.npz"), **head)

    # ---- forward operator goldens on the real series (L = 25,256) ----
    out = {}
    tracers = ["CFC12", "SF6", "H3", "He4_ter", "He3", "CFC11"]
    n = 24
    for name, (m1, m2, pn) in MODEL_CFGS.items():
        theta = draw_theta(pn, n, rng)
        out[name + "/theta"] = theta
        for t in tracers:
            out[name + "/" + t] = run_forward(ramu, ckw_real(C_in, t), m1, m2, pn, t, theta)
    np.savez_compressed(os.path.join(GOLD, "forward_real.npz"), **out)

    # ---- forward operator goldens on the synthetic monthly series (L = 840) ----
    series = synthetic.input_series(840, 0)
    out = {}
    n = 48
    for name in ("epm_dm", "emm0", "dm", "dm_dm", "epm_pfm123", "pfm"):
        m1, m2, pn = MODEL_CFGS[name]
        theta = draw_theta(pn, n, rng, tau_scale=12.0)
        out[name + "/theta"] = theta
        for t in synthetic.TRACERS_CFG3:
            out[name + "/" + t] = run_forward(ramu, ckw_synth(series, t), m1, m2, pn, t, theta)
    th3 = synthetic.theta_cfg3(64, 0)
    out["cfg3/theta"] = th3
    m1, m2, pn = MODEL_CFGS["epm_dm"]
    for t in synthetic.TRACERS_CFG3:
        out["cfg3/" + t] = run_forward(ramu, ckw_synth(series, t), m1, m2, list(synthetic.PAR_NAMES_CFG3), t, th3)
    th3i = synthetic.theta_cfg3_informative(64, 0)
    out["cfg3i/theta"] = th3i
    for t in synthetic.TRACERS_CFG3:
        out["cfg3i/" + t] = run_forward(ramu, ckw_synth(series, t), m1, m2, list(synthetic.PAR_NAMES_CFG3), t, th3i)
    np.savez_compressed(os.path.join(GOLD, "forward_synth.npz"), **out)

    # ---- gen_g_tp / convolve goldens (class API) ----
    out = {}
    c12 = C_in["CFC12"].iloc[-600:]
    for mt, kw in (("piston", dict(tau=10.5)), ("piston", dict(tau=0.3)), ("exponential", dict(tau=37.5)),
                   ("exp_pist_flow", dict(tau=37.5, eta=1.5)), ("exp_pist_flow", dict(tau=50.0, eta=1.0)),
                   ("dispersion", dict(tau=37.5, D=0.3)), ("dispersion", dict(tau=400.0, D=0.02)),
                   ("exp_pist_flow", dict(tau=1000.0, eta=5.0))):
        m = conv.tracer_conv_integral(c12.copy(), c12.index[-1])
        m.update_pars(mod_type=mt, t_half=20.0, **kw)
        with np.errstate(all="ignore"):
            g = m.gen_g_tp()
            c = m.convolve()
        key = mt + "_" + "_".join("%s%g" % kv for kv in kw.items())
        out[key + "/g"] = g
        out[key + "/C"] = np.array(c)
    # shifted sampling date (dtp = 3)
    m = conv.tracer_conv_integral(c12.copy(), c12.index[-1] + 3.2)
    m.update_pars(mod_type="exp_pist_flow", tau=25.0, eta=2.0, t_half=12.34)
    out["shift3/g"] = m.gen_g_tp()
    out["shift3/C"] = np.array(m.convolve())
    out["c12_600"] = c12.to_numpy().ravel()[::-1].copy()
    np.savez_compressed(os.path.join(GOLD, "rtd_weights.npz"), **out)

    # ---- CE model goldens ----
    gases = ["He", "Ne", "Ar", "Kr", "Xe"]
    n = 200
    E = rng.uniform(2700.0, 3300.0, n)
    T = rng.uniform(-1.0, 12.0, n)
    T[:6] = [70.0, 100.0, 99.0, 65.0, -0.5, 0.0]
    Ae = 10 ** rng.uniform(-4, -1, n)
    F = 10 ** rng.uniform(-1, 1, n)
    res = {k: np.empty((n, 5)) for k in ("ce_true", "ce_false", "eq_dry", "eq_wet", "K")}
    Pv = np.empty(n)
    Pl = np.empty(n)
    with np.errstate(all="ignore"):
        for i in range(n):
            o = ngu.noble_gas_fun(gases=gases, E=E[i], T=T[i], Ae=Ae[i], F=F[i], P="lapse_rate")
            a, b, c, d = o.ce_exc(True), o.ce_exc(False), o.equil_conc_dry(), o.equil_conc()
            for j, g in enumerate(gases):
                res["ce_true"][i, j], res["ce_false"][i, j] = a[g], b[g]
                res["eq_dry"][i, j], res["eq_wet"][i, j] = c[g], d[g]
                res["K"][i, j] = o.solubility(g)
            Pv[i], Pl[i] = o.vapor_pressure(), o.lapse_rate()
    np.savez_compressed(os.path.join(GOLD, "ce_model.npz"), E=E, T=T, Ae=Ae, F=F, P_vapor=Pv, P_lapse=Pl,
                        J_flux=np.array(ngu.J_flux(1, 2700, 1000, 3.7, 10.2, 0.05)), **res)
    # ---- fracture / matrix-diffusion RTD (numba-compiled in the reference), SURVEY 8f-4 ----
    out = {}
    c500 = C_in["CFC12"].iloc[-500:]
    for key, kw in (("a", dict(tau=30.0, D=0.3, bbar=1e-3, Phi_im=0.02)), ("b", dict(tau=120.0, D=0.05, bbar=5e-4, Phi_im=0.05)),
                    ("c", dict(tau=8.0, D=1.5, bbar=2e-3, Phi_im=0.01))):
        m = conv.tracer_conv_integral(c500.copy(), c500.index[-1])
        m.update_pars(mod_type="frac_inf_diff", t_half=25.0, **kw)
        with np.errstate(all="ignore"):
            out[key + "/g"] = m.gen_g_tp()
            out[key + "/FM_mu"] = np.array(m.FM_mu)
            out[key + "/C"] = np.array(m.convolve())
        out[key + "/par"] = np.array([kw["tau"], kw["D"], kw["bbar"], kw["Phi_im"]])
    out["c12_500"] = c500.to_numpy().ravel()[::-1].copy()
    # external advective RTD (frac_rtd_numba, :66-97): a gamma-shaped and a dispersion-shaped RTD on the yearly lag grid
    tpe = np.arange(500, dtype=float)
    ext = {"d": tpe * np.exp(-tpe / 15.0), "e": np.where(tpe > 0, np.exp(-(1 - np.maximum(tpe, 1e-9) / 40.0) ** 2 / (4 * 0.2 * np.maximum(tpe, 1e-9) / 40.0))
                                                         / np.maximum(tpe, 1e-9) ** 1.5, 0.0)}
    for key, kw in (("d", dict(bbar=1e-3, Phi_im=0.02)), ("e", dict(bbar=5e-4, Phi_im=0.05))):
        fe = ext[key] / ext[key].sum()
        m = conv.tracer_conv_integral(c500.copy(), c500.index[-1])
        m.update_pars(mod_type="frac_inf_diff", t_half=25.0, f_tadv_ext=fe, **kw)
        with np.errstate(all="ignore"):
            out[key + "/g"] = m.gen_g_tp()
            out[key + "/FM_mu"] = np.array(m.FM_mu)
            out[key + "/C"] = np.array(m.convolve())
        out[key + "/par"] = np.array([kw["bbar"], kw["Phi_im"]])
        out[key + "/f_tadv_ext"] = fe
    np.savez_compressed(os.path.join(GOLD, "fdm_weights.npz"), **out)

    # ---- CFC / SF6 corrections (utils/cfc_utils.py), SURVEY 8f-2 ----
    import cfc_utils as ref_cfc
    n = 120
    E = rng.uniform(2700.0, 3300.0, n); T = rng.uniform(0.1, 12.0, n); Ae = 10 ** rng.uniform(-4, -1, n); F = 10 ** rng.uniform(-1, 0.5, n)
    Cm = rng.uniform(0.2, 5.0, (n, 3)); zi = rng.uniform(50.0, 550.0, (n, 3)); Cs = rng.uniform(0.1, 3.0, n); zs = rng.uniform(1.0, 11.0, n)
    r = {k: np.empty((n, 3)) for k in ("cfc_air", "cfc_aq", "cfc_exc", "cfc_K")}
    q = {k: np.empty(n) for k in ("sf6_air", "sf6_aq", "sf6_exc", "sf6_K")}
    for i in range(n):
        c = ref_cfc.cfc_ce_corr(cfc_num=[11, 12, 113], E=E[i], T=T[i], Ae=Ae[i], F=F[i])
        r["cfc_K"][i] = c.solubility_cfc(); r["cfc_air"][i] = c.equil_air_conc_cfc(Cm[i]); r["cfc_aq"][i] = c.equil_aq_conc_cfc(zi[i])
        r["cfc_exc"][i] = c.ce_exc_conc_cfc(zi[i])
        s6 = ref_cfc.sf6_ce_corr(E=E[i], T=T[i], Ae=Ae[i], F=F[i])
        q["sf6_K"][i] = s6.solubility_sf6(); q["sf6_air"][i] = s6.equil_air_conc_sf6(Cs[i]); q["sf6_aq"][i] = s6.equil_aq_conc_sf6(zs[i])
        q["sf6_exc"][i] = s6.ce_exc_conc_sf6(zs[i])
    np.savez_compressed(os.path.join(GOLD, "cfc_model.npz"), E=E, T=T, Ae=Ae, F=F, Cm=Cm, zi=zi, Cs=Cs, zs=zs, **r, **q)

    # ---- posterior known-answer fixture for config 1 (the only posterior summaries the reference ships) ----
    import csv
    import json
    ng_dir = os.path.join(ref_shims.REFERENCE_ROOT, "ng_interp")
    post = {"source": "ng_interp/ng_optPLM{1,6,7}.csv (az.summary + median) and panga_comp/ng_conc_4_panga.csv of the reference",
            "wells": {}}
    obs = {}
    with open(os.path.join(ng_dir, "panga_comp", "ng_conc_4_panga.csv")) as f:
        for row in csv.DictReader(f):
            obs[row["wells"]] = {g: float(row[g]) for g in ("He", "Ne", "Ar", "Kr", "Xe")}
    for w in ("PLM1", "PLM6", "PLM7"):
        with open(os.path.join(ng_dir, "ng_opt%s.csv" % w)) as f:
            rows = list(csv.reader(f))
        hdr = rows[0][1:]
        post["wells"][w] = {"obs": obs[w], "summary": {r[0]: dict(zip(hdr, map(float, r[1:]))) for r in rows[1:] if r}}
    with open(os.path.join(GOLD, "ng_posterior.json"), "w") as f:
        json.dump(post, f, indent=1)

    print("golden vectors written to", GOLD)
    for f in sorted(os.listdir(GOLD)):
        print("  %-24s %8d bytes" % (f, os.path.getsize(os.path.join(GOLD, f))))


if __name__ == "__main__":
    main()
