"""CPU: pin the oracle restatement (oracle/np_oracle.py) against vectors produced by the untouched
reference (tests/golden/*, made by oracle/gen_golden.py) and the SURVEY App. C known answers."""
import os

import numpy as np
import pytest

import np_oracle as O
from helpers import GOLD, J_MU, MODEL_CFGS, REAL_TRACERS, load_c_in, rel_err
from noblegas_rtd_mcmc_b200 import synthetic

TOL = 1e-12


def _fwd_real(C, name, tracer, theta):
    m1, m2, pn = MODEL_CFGS[name]
    s, th, ra = REAL_TRACERS[tracer]
    series = C[s] if s is not None else np.zeros(len(C["H3"]))
    return O.forward_mod(theta, pn, tracer, series, m1, m2, t_half=th, rad_accum=ra)


@pytest.mark.parametrize("name", sorted(MODEL_CFGS))
def test_forward_real_series(name):
    z = np.load(os.path.join(GOLD, "forward_real.npz"))
    C = load_c_in()
    theta = z[name + "/theta"]
    for tracer in ("CFC12", "SF6", "H3", "He4_ter", "He3", "CFC11"):
        got = _fwd_real(C, name, tracer, theta)
        assert rel_err(got, z[name + "/" + tracer]) < TOL, (name, tracer)


@pytest.mark.parametrize("name", ["epm_dm", "emm0", "dm", "dm_dm", "epm_pfm123", "pfm", "cfg3", "cfg3i"])
def test_forward_synthetic_monthly(name):
    z = np.load(os.path.join(GOLD, "forward_synth.npz"))
    series = synthetic.input_series(840, 0)
    tab = synthetic.tracer_table_cfg3()
    if name.startswith("cfg3"):
        m1, m2, pn = "exp_pist_flow", "dispersion", list(synthetic.PAR_NAMES_CFG3)
    else:
        m1, m2, pn = MODEL_CFGS[name]
    theta = z[name + "/theta"]
    for tracer in synthetic.TRACERS_CFG3:
        d = tab[tracer]
        s = series[d["series"]] if d["series"] is not None else np.zeros(840)
        got = O.forward_mod(theta, pn, tracer, s, m1, m2, t_half=d.get("t_half", False),
                            rad_accum=d.get("rad_accum", False))
        assert rel_err(got, z[name + "/" + tracer]) < TOL, (name, tracer)


def test_cfg3_theta_is_reproducible():
    z = np.load(os.path.join(GOLD, "forward_synth.npz"))
    assert np.array_equal(z["cfg3/theta"], synthetic.theta_cfg3(64, 0))
    assert np.array_equal(z["cfg3i/theta"], synthetic.theta_cfg3_informative(64, 0))


def test_survey_appendix_c_known_answers():
    """values printed by the reference for tau=37.5 (eta=1.5, D=0.3)."""
    C = load_c_in()
    J = O.J_flux(1, 2700, 1000, 3.0, 10.0, 0.05)
    assert J == 3.308849999999999e-11
    tp = O.lag_grid(len(C["H3"]))
    table = {
        ("H3", "exponential"): 5.769061780670842, ("H3", "piston"): 2.711377625858228,
        ("H3", "exp_pist_flow"): 5.735133553732442, ("H3", "dispersion"): 5.683112393627562,
        ("CFC12", "exponential"): 357.84811996405404, ("CFC12", "piston"): 359.4401250000999,
        ("CFC12", "exp_pist_flow"): 367.0441509423791, ("CFC12", "dispersion"): 367.3215610361376,
        ("SF6", "exponential"): 3.8897555359721503, ("SF6", "piston"): 1.2233333334333334,
        ("SF6", "exp_pist_flow"): 2.6236937909320046, ("SF6", "dispersion"): 2.9017305223555776,
        ("He4_ter", "exponential"): 1.2243480377199597e-09, ("He4_ter", "piston"): 1.2242744999999997e-09,
        ("He4_ter", "exp_pist_flow"): 1.2409290420589117e-09, ("He4_ter", "dispersion"): 1.2408187310121764e-09,
        ("He3", "exponential"): 66.89644005427098,
    }
    for (tracer, mt), want in table.items():
        s, th, ra = REAL_TRACERS[tracer]
        series = C[s] if s is not None else np.zeros(len(tp))
        g = O.gen_g_tp(mt, tp, 37.5, eta=1.5, D=0.3)
        lam = O.thalf_2_lambda(th) if th else 0.0
        got = O.convolve(series, tp, g, lam, ra, J)[0]
        assert abs(got - want) <= 1e-12 * abs(want), (tracer, mt, got, want)
    # operator boundary
    theta = np.array([[42.0, 1.7, J_MU + 0.1, 20.0, 0.05]])
    pn = ["tau1", "eta1", "J", "thalf_cfc", "lamsf6"]
    want = {"CFC12": 127.95406693711098, "SF6": 2.0408601071408903, "H3": 6.151643687622419,
            "He4_ter": 2.0010438892179075e-09, "He3": 97.14969224920476}
    for tracer, w in want.items():
        s, th, ra = REAL_TRACERS[tracer]
        series = C[s] if s is not None else np.zeros(len(tp))
        got = O.forward_mod(theta, pn, tracer, series, "exp_pist_flow", False, t_half=th, rad_accum=ra)[0]
        assert abs(got - w) <= 1e-12 * abs(w), (tracer, got, w)
    # edges
    g = O.gen_g_tp("piston", tp, 10.5)
    assert g[0].argmax() == 10
    assert abs(O.convolve(C["CFC12"], tp, g)[0] - 532.8290000001001) < 1e-9
    a = O.convolve(C["CFC12"], tp, O.gen_g_tp("exp_pist_flow", tp, 50.0, eta=1.0))[0]
    b = O.convolve(C["CFC12"], tp, O.gen_g_tp("exponential", tp, 50.0))[0]
    assert a == b and abs(a - 303.8326702342745) < 1e-10
    assert O.convolve(C["CFC12"], tp, O.gen_g_tp("exp_pist_flow", tp, 1000.0, eta=5.0))[0] == pytest.approx(1e-10, rel=1e-9)


def test_rtd_weights_and_shift():
    z = np.load(os.path.join(GOLD, "rtd_weights.npz"))
    c = z["c12_600"]
    tp = O.lag_grid(600)
    lam = O.thalf_2_lambda(20.0)
    cases = {"piston_tau10.5": ("piston", dict(tau=10.5)), "piston_tau0.3": ("piston", dict(tau=0.3)),
             "exponential_tau37.5": ("exponential", dict(tau=37.5)),
             "exp_pist_flow_tau37.5_eta1.5": ("exp_pist_flow", dict(tau=37.5, eta=1.5)),
             "exp_pist_flow_tau50_eta1": ("exp_pist_flow", dict(tau=50.0, eta=1.0)),
             "dispersion_tau37.5_D0.3": ("dispersion", dict(tau=37.5, D=0.3)),
             "dispersion_tau400_D0.02": ("dispersion", dict(tau=400.0, D=0.02)),
             "exp_pist_flow_tau1000_eta5": ("exp_pist_flow", dict(tau=1000.0, eta=5.0))}
    for key, (mt, kw) in cases.items():
        g = O.gen_g_tp(mt, tp, kw["tau"], eta=kw.get("eta"), D=kw.get("D"))
        assert rel_err(g[0], z[key + "/g"]) < 1e-13, key
        assert rel_err(O.convolve(c, tp, g, lam), z[key + "/C"].reshape(1)) < 1e-12, key
    tp3 = O.lag_grid(600, 3.0)
    g = O.gen_g_tp("exp_pist_flow", tp3, 25.0, eta=2.0)
    assert rel_err(g[0], z["shift3/g"]) < 1e-13
    assert rel_err(O.convolve(c, tp3, g, O.thalf_2_lambda(12.34)), z["shift3/C"].reshape(1)) < 1e-12


def test_ce_model_golden():
    z = np.load(os.path.join(GOLD, "ce_model.npz"))
    gases = ["He", "Ne", "Ar", "Kr", "Xe"]
    E, T, Ae, F = z["E"], z["T"], z["Ae"], z["F"]
    assert rel_err(O.lapse_rate(E), z["P_lapse"]) < 1e-14
    assert rel_err(O.vapor_pressure(T), z["P_vapor"]) < 1e-14
    K = np.stack([O.solubility(g, T) for g in gases], axis=1)
    assert rel_err(K, z["K"]) < 1e-14
    assert rel_err(O.equil_conc_dry(gases, T, O.lapse_rate(E)), z["eq_dry"]) < 1e-14
    assert rel_err(O.equil_conc(gases, T, O.lapse_rate(E)), z["eq_wet"]) < 1e-14
    assert rel_err(O.ce_exc(gases, E, T, Ae, F, True), z["ce_true"]) < 1e-14
    assert rel_err(O.ce_exc(gases, E, T, Ae, F, False), z["ce_false"]) < 1e-14
    assert float(z["J_flux"]) == O.J_flux(1, 2700, 1000, 3.7, 10.2, 0.05) == 3.7657277999999995e-11


def test_ce_known_answers_appendix_c():
    gases = ["He", "Ne", "Ar", "Kr", "Xe"]
    got = O.ce_exc(gases, 3000.0, 3.5, 0.01, 0.5, True)[0]
    want = [4.885128121481438e-08, 2.0811831959479197e-07, 0.00036126939191904594, 8.376899405441966e-08,
            1.2509731676800307e-08]
    assert np.allclose(got, want, rtol=1e-14, atol=0)
    neg = O.ce_exc(gases, 2900.0, -0.5, 0.01, 0.5, True)[0]
    assert np.allclose(neg, [-9998.9999999738, -9998.9999999091, -9998.9999533, -9998.9999999943, -9998.999999999565],
                       rtol=1e-12)
    w = O.ce_exc_wrapper(np.array([[np.log10(0.019898), np.log10(0.390867), 2974.177443, 1.325121]]))[0]
    assert np.allclose(w, [2.720123026243957e-07, 0.00043618492861502585, 9.755707127645672e-08, 1.4430312435059888e-08],
                       rtol=1e-13)


def test_logp_against_scipy():
    """pymc3 is absent (third party); the likelihood formulae are pinned against scipy.stats."""
    from scipy import stats
    rng = np.random.default_rng(3)
    mu = rng.normal(1.0, 0.5, (50, 4))
    obs = np.array([1.1, 0.7, 1.3, 0.9])
    sd = np.array([0.1, 0.2, 0.05, 0.3])
    nu = rng.uniform(1.0, 30.0, 50)
    want_n = stats.norm.logpdf(obs, mu, sd).sum(axis=1)
    want_t = stats.t.logpdf(obs, nu[:, None], mu, sd).sum(axis=1)
    assert np.allclose(O.logp_normal(obs, mu, sd), want_n, rtol=1e-13)
    assert np.allclose(O.logp_studentt(obs, mu, sd, nu), want_t, rtol=1e-12)


@pytest.mark.needs_reference
def test_oracle_matches_live_reference():
    """Build container only: re-run the untouched reference on fresh draws and compare."""
    import pandas as pd
    import ref_shims
    conv, ngu, ramu = ref_shims.load()
    rng = np.random.default_rng(7)
    L = 400
    series = rng.uniform(0.1, 10.0, L)
    df = pd.DataFrame({"X": series[::-1]}, index=np.arange(L - 1, -1, -1))
    tp = O.lag_grid(L)
    for mt in O.MOD_TYPES:
        tau, eta, D = rng.uniform(2, 300), rng.uniform(1, 5), rng.uniform(0.01, 2)
        m = conv.tracer_conv_integral(df.copy(), 0)
        m.update_pars(tau=tau, mod_type=mt, eta=eta, D=D, t_half=9.0)
        want = m.convolve()
        got = O.convolve(series, tp, O.gen_g_tp(mt, tp, tau, eta, D), O.thalf_2_lambda(9.0))[0]
        assert abs(got - want) <= 1e-13 * abs(want), mt


# ----------------------------------------------------------------------------- plain-C oracle
def _real_descs(tracers, par_names):
    names = ["CFC11", "CFC12", "CFC113", "SF6", "H3"]
    descs = []
    for t in tracers:
        s, th, ra = REAL_TRACERS[t]
        descs.append(dict(series=names.index(s) if s is not None else -1, rad_accum=ra,
                          lam=float(-1.0 * np.log(0.5) / th) if th else 0.0,
                          use_thalf_cfc=(t == "CFC12" and "thalf_cfc" in par_names), use_lamsf6=(t == "SF6")))
    return names, descs


@pytest.mark.parametrize("name", sorted(MODEL_CFGS))
def test_c_oracle_forward_real_series(name):
    import c_oracle
    z = np.load(os.path.join(GOLD, "forward_real.npz"))
    C = load_c_in()
    m1, m2, pn = MODEL_CFGS[name]
    tracers = ["CFC12", "SF6", "H3", "He4_ter", "He3", "CFC11"]
    names, descs = _real_descs(tracers, pn)
    X = np.stack([C[n] for n in names], axis=1)
    out = c_oracle.forward(X, descs, m1, m2, z[name + "/theta"], pn)
    for i, t in enumerate(tracers):
        assert rel_err(out[:, i], z[name + "/" + t]) < 1e-11, (name, t)


def test_c_oracle_ce_and_loglik():
    import c_oracle
    z = np.load(os.path.join(GOLD, "ce_model.npz"))
    gases = ["He", "Ne", "Ar", "Kr", "Xe"]
    for what, key in ((0, "ce_true"), (1, "ce_false"), (2, "eq_dry"), (3, "eq_wet"), (4, "K")):
        got = c_oracle.ce(what, gases, z["E"], z["T"], z["Ae"], z["F"])
        assert rel_err(got, z[key]) < 1e-13, key
    rng = np.random.default_rng(3)
    mu = rng.normal(1.0, 0.5, (50, 4))
    obs = np.array([1.1, 0.7, 1.3, 0.9]); sd = np.array([0.1, 0.2, 0.05, 0.3]); nu = rng.uniform(1.0, 30.0, 50)
    assert np.allclose(c_oracle.loglik("normal", mu, obs, sd), O.logp_normal(obs, mu, sd), rtol=1e-13)
    assert np.allclose(c_oracle.loglik("studentt", mu, obs, sd, nu), O.logp_studentt(obs, mu, sd, nu), rtol=1e-12)


def test_cfc_sf6_corrections_golden():
    """utils/cfc_utils.py: numpy restatement vs the untouched reference."""
    z = np.load(os.path.join(GOLD, "cfc_model.npz"))
    E, T, Ae, F = z["E"], z["T"], z["Ae"], z["F"]
    for what, key, X in (("K", "cfc_K", None), ("air", "cfc_air", z["Cm"]), ("aq", "cfc_aq", z["zi"]), ("exc", "cfc_exc", z["zi"])):
        assert rel_err(O.cfc_corr(what, [11, 12, 113], E, T, Ae, F, X), z[key]) < 1e-13, key
    for what, key, X in (("K", "sf6_K", None), ("air", "sf6_air", z["Cs"]), ("aq", "sf6_aq", z["zs"]), ("exc", "sf6_exc", z["zs"])):
        Xa = None if X is None else X.reshape(-1, 1)
        assert rel_err(O.cfc_corr(what, [6], E, T, Ae, F, Xa)[:, 0], z[key]) < 1e-13, key


def test_frac_inf_diff_weights_golden():
    """the numba-compiled fracture / matrix-diffusion RTD of the reference vs the numpy restatement,
    including the SURVEY known answer (CFC-12 newest 500 rows, tau=30, D=0.3, bbar=1e-3, Phi_im=0.02)."""
    z = np.load(os.path.join(GOLD, "fdm_weights.npz"))
    tp = O.lag_grid(500)
    for k in "abc":
        tau, D, bbar, phi = z[k + "/par"]
        g, mu = O.frac_inf_diff_weights(tp, tau, D, bbar, phi)
        assert rel_err(g, z[k + "/g"]) < 1e-12, k
        assert abs(mu - float(z[k + "/FM_mu"])) < 1e-12 * mu
        C = O.convolve(z["c12_500"], tp, g.reshape(1, -1), O.thalf_2_lambda(25.0))[0]
        assert abs(C - float(z[k + "/C"])) < 1e-12 * abs(C)
    assert abs(float(z["a/FM_mu"]) - 86.31699496976702) < 1e-11
    for k in "de":                                          # caller-supplied advective RTD (frac_rtd_numba, :66-97)
        bbar, phi = z[k + "/par"]
        g, mu = O.frac_inf_diff_weights(tp, None, None, bbar, phi, f_tadv_ext=z[k + "/f_tadv_ext"])
        assert rel_err(g, z[k + "/g"]) < 1e-12, k
        assert abs(mu - float(z[k + "/FM_mu"])) < 1e-12 * mu
        C = O.convolve(z["c12_500"], tp, g.reshape(1, -1), O.thalf_2_lambda(25.0))[0]
        assert abs(C - float(z[k + "/C"])) < 1e-12 * abs(C)
