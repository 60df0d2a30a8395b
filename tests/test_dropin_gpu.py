"""GPU: the drop-in Python classes (reference call signatures) against golden vectors of the untouched reference."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
TOL = 1e-10


def _df(values_newest_first, name="X"):
    import pandas as pd
    L = len(values_newest_first)
    return pd.DataFrame({name: np.asarray(values_newest_first)[::-1]}, index=np.arange(L - 1, -1, -1))


def test_tracer_conv_integral_weights_and_convolve():
    from helpers import GOLD, rel_err
    from noblegas_rtd_mcmc_b200.convolution_integral_utils import tracer_conv_integral
    z = np.load(os.path.join(GOLD, "rtd_weights.npz"))
    c12 = _df(z["c12_600"], "CFC12")
    cases = {"piston_tau10.5": ("piston", dict(tau=10.5)), "piston_tau0.3": ("piston", dict(tau=0.3)),
             "exponential_tau37.5": ("exponential", dict(tau=37.5)),
             "exp_pist_flow_tau37.5_eta1.5": ("exp_pist_flow", dict(tau=37.5, eta=1.5)),
             "exp_pist_flow_tau50_eta1": ("exp_pist_flow", dict(tau=50.0, eta=1.0)),
             "dispersion_tau37.5_D0.3": ("dispersion", dict(tau=37.5, D=0.3)),
             "dispersion_tau400_D0.02": ("dispersion", dict(tau=400.0, D=0.02)),
             "exp_pist_flow_tau1000_eta5": ("exp_pist_flow", dict(tau=1000.0, eta=5.0))}
    for key, (mt, kw) in cases.items():
        m = tracer_conv_integral(c12.copy(), c12.index[-1])
        m.update_pars(mod_type=mt, t_half=20.0, **kw)
        g = m.gen_g_tp()
        assert g.shape == (600,)
        assert rel_err(g, z[key + "/g"]) < 1e-12, key
        C = m.convolve()
        assert isinstance(C, float)
        assert rel_err(np.array([C]), z[key + "/C"].reshape(1)) < TOL, key
        # external weights path: convolve(g_tau=...)
        C2 = m.convolve(g_tau=z[key + "/g"])
        assert rel_err(np.array([C2]), z[key + "/C"].reshape(1)) < 1e-12, key
    m = tracer_conv_integral(c12.copy(), c12.index[-1] + 3.2)       # shifted sampling date, dtp = 3
    m.update_pars(mod_type="exp_pist_flow", tau=25.0, eta=2.0, t_half=12.34)
    assert rel_err(m.gen_g_tp(), z["shift3/g"]) < 1e-12
    assert rel_err(np.array([m.convolve()]), z["shift3/C"].reshape(1)) < TOL


def test_tracer_conv_integral_appendix_c_and_attribute_pokes():
    """SURVEY App. C values through the class API, including the posterior-predictive idiom of the reference
    (attributes poked between convolve() calls, thalf_2_lambda after update_pars: run_age_mcmc.py:293-303)."""
    from helpers import load_c_in
    from noblegas_rtd_mcmc_b200.convolution_integral_utils import tracer_conv_integral
    from noblegas_rtd_mcmc_b200.noble_gas_utils import J_flux
    C = load_c_in()
    m = tracer_conv_integral(_df(C["CFC12"], "CFC12"), 0)
    m.update_pars(tau=37.5, mod_type="exponential")
    assert abs(m.convolve() - 357.84811996405404) < 1e-10 * 357.8
    m.mod_type, m.eta = "exp_pist_flow", 1.5
    assert abs(m.convolve() - 367.0441509423791) < 1e-10 * 367.0
    m.mod_type, m.D = "dispersion", 0.3
    assert abs(m.convolve() - 367.3215610361376) < 1e-10 * 367.3
    m.mod_type, m.tau = "piston", 10.5
    assert abs(m.convolve() - 532.8290000001001) < 1e-9
    h = tracer_conv_integral(_df(C["H3"], "H3_tu"), 0)
    h.update_pars(tau=37.5, mod_type="exponential", t_half=12.34, rad_accum="3He")
    assert abs(h.convolve() - 66.89644005427098) < 1e-10 * 66.9
    h.update_pars(tau=37.5, mod_type="exponential")
    h.thalf_2_lambda(12.34)
    assert abs(h.convolve() - 5.769061780670842) < 1e-10 * 5.77
    he = tracer_conv_integral(_df(np.zeros(len(C["H3"])), "He4_ter"), 0)
    he.update_pars(tau=37.5, mod_type="dispersion", D=0.3, rad_accum="4He", J=J_flux(1, 2700, 1000, 3.0, 10.0, 0.05))
    assert abs(he.convolve() - 1.2408187310121764e-09) < 1e-10 * 1.24e-9
    # batched (array-valued tau) is the additive behaviour
    he.tau = np.array([37.5, 100.0, 400.0])
    out = he.convolve()
    assert out.shape == (3,) and abs(out[0] - 1.2408187310121764e-09) < 1e-10 * 1.24e-9
    with pytest.raises(ValueError):
        he.mod_type = "gamma"
        he.convolve()


def test_noble_gas_fun_golden():
    from helpers import GOLD, rel_err
    from noblegas_rtd_mcmc_b200.noble_gas_utils import J_flux, atm_std, noble_gas_fun
    z = np.load(os.path.join(GOLD, "ce_model.npz"))
    gases = ["He", "Ne", "Ar", "Kr", "Xe"]
    ng = noble_gas_fun(gases=gases, E=z["E"], T=z["T"], Ae=z["Ae"], F=z["F"], P="lapse_rate")
    for meth, arg, key in (("ce_exc", (True,), "ce_true"), ("ce_exc", (False,), "ce_false"),
                           ("equil_conc_dry", (), "eq_dry"), ("equil_conc", (), "eq_wet")):
        d = getattr(ng, meth)(*arg)
        got = np.stack([d[g] for g in gases], axis=1)
        assert rel_err(got, z[key]) < 1e-12, key
    K = np.stack([ng.solubility(g) for g in gases], axis=1)
    assert rel_err(K, z["K"]) < 1e-12
    assert rel_err(ng.vapor_pressure(), z["P_vapor"]) < 1e-13 and rel_err(ng.lapse_rate(), z["P_lapse"]) < 1e-13
    # scalar call: plain floats keyed by gas, App. C known answer, sentinel for T < 0
    s = noble_gas_fun(gases=gases, E=3000.0, T=3.5, Ae=0.01, F=0.5, P="lapse_rate").ce_exc(True)
    want = [4.885128121481438e-08, 2.0811831959479197e-07, 0.00036126939191904594, 8.376899405441966e-08,
            1.2509731676800307e-08]
    assert all(isinstance(s[g], float) for g in gases)
    assert np.allclose([s[g] for g in gases], want, rtol=1e-12, atol=0)
    neg = noble_gas_fun(gases=gases, E=2900.0, T=-0.5, Ae=0.01, F=0.5, P="lapse_rate")
    assert all(v == -9999.0 for v in neg.equil_conc_dry().values())
    assert atm_std["Ne"] == 1.818e-5 and J_flux(1, 2700, 1000, 3.7, 10.2, 0.05) == 3.7657277999999995e-11
    ng.update_pars(T=3.5, Ae=0.01, F=0.5, E=3000.0)
    assert abs(ng.ce_exc(True)["Ar"] - want[2]) < 1e-12 * want[2]


def test_forwardmod_operator_and_ce_wrapper():
    from helpers import J_MU, load_c_in
    from noblegas_rtd_mcmc_b200.noble_gas_mcmc import ce_exc_wrapper
    from noblegas_rtd_mcmc_b200.run_age_mcmc_utils import ForwardMod, JointForwardMod
    C = load_c_in()
    ckw = {"CFC12": dict(C_t=_df(C["CFC12"], "CFC12")), "SF6": dict(C_t=_df(C["SF6"], "SF6")),
           "H3": dict(C_t=_df(C["H3"], "H3_tu"), t_half=12.34),
           "He4_ter": dict(C_t=_df(C["He4_ter"], "He4_ter") * 0.0, rad_accum="4He"),
           "He3": dict(C_t=_df(C["H3"], "H3_tu"), t_half=12.34, rad_accum="3He")}
    for kw in ckw.values():
        kw["mod_type1"], kw["mod_type2"] = "exp_pist_flow", False
    pn = ["tau1", "eta1", "J", "thalf_cfc", "lamsf6"]
    theta = np.array([42.0, 1.7, J_MU + 0.1, 20.0, 0.05])
    want = {"CFC12": 127.95406693711098, "SF6": 2.0408601071408903, "H3": 6.151643687622419,
            "He4_ter": 2.0010438892179075e-09, "He3": 97.14969224920476}
    for t, w in want.items():
        op = ForwardMod(ckw[t], pn, t)
        out = [[None]]
        op.perform(None, [theta], out)
        assert out[0][0].shape == () and abs(float(out[0][0]) - w) < TOL * abs(w), t
    joint = JointForwardMod(ckw, pn, list(want))
    res = joint.perform_batch(np.stack([theta, theta]))
    assert res.shape == (2, 5)
    assert np.allclose(res[1], list(want.values()), rtol=TOL, atol=0)
    # EPM + PFM mixture, App. C
    for kw in ckw.values():
        kw["mod_type2"] = "piston"
    pn2 = ["tau1", "tau2", "f1", "f2", "eta1", "J", "thalf_cfc", "lamsf6"]
    th2 = np.array([[30, 2500, 0.6, 0.4, 2.2, J_MU, 15, 0.1]])
    got = JointForwardMod(ckw, pn2, ["CFC12", "SF6", "H3", "He4_ter"]).perform_batch(th2)[0]
    assert np.allclose(got, [85.59104057853781, 1.9587812675737937, 2.8588643900529678, 3.8338328118329954e-08], rtol=TOL)
    w = ce_exc_wrapper(np.array([np.log10(0.019898), np.log10(0.390867), 2974.177443, 1.325121]))
    assert np.allclose(w, [2.720123026243957e-07, 0.00043618492861502585, 9.755707127645672e-08, 1.4430312435059888e-08],
                       rtol=1e-12)
    with pytest.raises(ValueError):
        JointForwardMod(ckw, ["tau1", "bogus"], ["CFC12"])


def test_fused_loglik_matches_oracle():
    import np_oracle as O
    from helpers import synth_plan
    from noblegas_rtd_mcmc_b200 import synthetic
    pn = list(synthetic.PAR_NAMES_CFG3)
    plan, _, _ = synth_plan("exp_pist_flow", "dispersion", pn)
    theta = synthetic.theta_cfg3_informative(1000, 11)
    model = plan.forward_host(theta, pn)
    obs = np.nanmedian(model, axis=0)
    sd = 0.05 * np.abs(obs)
    nu = np.random.default_rng(1).uniform(5, 30, 1000)
    lp_n, model2 = plan.forward_loglik_host(theta, pn, obs, sd, "normal", want_model=True)
    lp_t = plan.forward_loglik_host(theta, pn, obs, sd, "studentt", nu=nu)
    assert np.array_equal(model, model2)
    assert np.allclose(lp_n, O.logp_normal(obs, model, sd), rtol=1e-12, equal_nan=True)
    assert np.allclose(lp_t, O.logp_studentt(obs, model, sd, nu), rtol=1e-12, equal_nan=True)


def test_ng_parse_he_comps_batched():
    """He_comps over a batch of CE parameter draws against the oracle arithmetic."""
    import np_oracle as O
    from helpers import GOLD
    from noblegas_rtd_mcmc_b200.noble_gas_utils import ng_parse
    z = np.load(os.path.join(GOLD, "ce_model.npz"))
    sel = z["T"] > 0.0
    E, T, Ae, F = z["E"][sel], z["T"][sel], z["Ae"][sel], z["F"][sel]
    obs = {"He4": 7.283e-08, "He3": 1.1e-13}
    p = ng_parse(obs, Ae, F, E, T)
    p.He_comps(Rterr=2.0e-8)
    eq = O.equil_conc(["He"], T, O.lapse_rate(E))[:, 0]
    atm = O.ce_exc(["He"], E, T, Ae, F, True)[:, 0]
    ter = obs["He4"] - atm
    tu = (obs["He3"] - (obs["He4"] - ter) * 1.384e-6 + eq * 1.384e-6 * (1 - 0.983) - ter * 2.0e-8) * 4.021e14
    assert np.allclose(p.obs_dict_["He4_eq"], eq, rtol=1e-12) and np.allclose(p.obs_dict_["He4_ter"], ter, rtol=1e-9, atol=1e-22)
    assert np.allclose(p.obs_dict_["He3_tu"], tu, rtol=1e-9, atol=1e-9)
    s = ng_parse(obs, 0.01, 0.5, 3000.0, 3.5)
    s.He_comps(2.0e-8)
    assert isinstance(s.obs_dict_["He4_ter"], float)


def test_cfc_sf6_dropin_golden():
    """cfc_ce_corr / sf6_ce_corr drop-ins (batched and scalar) against the untouched reference."""
    from helpers import GOLD, rel_err
    from noblegas_rtd_mcmc_b200.cfc_utils import cfc_ce_corr, sf6_ce_corr
    z = np.load(os.path.join(GOLD, "cfc_model.npz"))
    E, T, Ae, F = z["E"], z["T"], z["Ae"], z["F"]
    c = cfc_ce_corr(cfc_num=[11, 12, 113], E=E, T=T, Ae=Ae, F=F)
    assert rel_err(c.solubility_cfc(), z["cfc_K"]) < 1e-12
    assert rel_err(c.equil_air_conc_cfc(z["Cm"]), z["cfc_air"]) < 1e-12
    assert rel_err(c.equil_aq_conc_cfc(z["zi"]), z["cfc_aq"]) < 1e-12
    assert rel_err(c.ce_exc_conc_cfc(z["zi"]), z["cfc_exc"]) < 1e-12
    s6 = sf6_ce_corr(E=E, T=T, Ae=Ae, F=F)
    assert rel_err(s6.solubility_sf6(), z["sf6_K"]) < 1e-12
    assert rel_err(s6.equil_air_conc_sf6(z["Cs"]), z["sf6_air"]) < 1e-12
    assert rel_err(s6.equil_aq_conc_sf6(z["zs"]), z["sf6_aq"]) < 1e-12
    assert rel_err(s6.ce_exc_conc_sf6(z["zs"]), z["sf6_exc"]) < 1e-12
    one = cfc_ce_corr(cfc_num=[11, 12, 113], E=float(E[3]), T=float(T[3]), Ae=float(Ae[3]), F=float(F[3]))
    assert one.equil_air_conc_cfc(z["Cm"][3]).shape == (3,) and rel_err(one.equil_air_conc_cfc(z["Cm"][3]), z["cfc_air"][3]) < 1e-12
    v = sf6_ce_corr(E=float(E[3]), T=float(T[3]), Ae=float(Ae[3]), F=float(F[3])).equil_air_conc_sf6(float(z["Cs"][3]))
    assert isinstance(v, float) and abs(v - z["sf6_air"][3]) < 1e-12 * abs(v)
    with pytest.raises(ValueError):
        cfc_ce_corr(cfc_num=[13], E=1.0, T=1.0, Ae=0.1, F=0.1)


def test_frac_inf_diff_dropin_golden():
    """tracer_conv_integral(mod_type='frac_inf_diff') on the GPU vs the reference's numba implementation."""
    from helpers import GOLD, rel_err
    from noblegas_rtd_mcmc_b200.convolution_integral_utils import tracer_conv_integral
    z = np.load(os.path.join(GOLD, "fdm_weights.npz"))
    c12 = _df(z["c12_500"], "CFC12")
    for k in "abc":
        tau, D, bbar, phi = z[k + "/par"]
        m = tracer_conv_integral(c12.copy(), c12.index[-1])
        m.update_pars(mod_type="frac_inf_diff", t_half=25.0, tau=tau, D=D, bbar=bbar, Phi_im=phi)
        g = m.gen_g_tp()
        assert rel_err(g, z[k + "/g"]) < 1e-10, k
        assert abs(m.FM_mu - float(z[k + "/FM_mu"])) < 1e-10 * m.FM_mu
        assert abs(m.convolve() - float(z[k + "/C"])) < 1e-10 * abs(float(z[k + "/C"]))
    m.tau = np.array([30.0, 120.0])                       # batched parameters
    m.D, m.bbar, m.Phi_im = np.array([0.3, 0.05]), np.array([1e-3, 5e-4]), np.array([0.02, 0.05])
    out = m.convolve()
    assert out.shape == (2,) and abs(out[0] - float(z["a/C"])) < 1e-10 * out[0] and abs(out[1] - float(z["b/C"])) < 1e-10 * out[1]
    for k in "de":                                        # caller-supplied advective RTD (frac_rtd_numba, :66-97)
        bbar, phi = z[k + "/par"]
        m = tracer_conv_integral(c12.copy(), c12.index[-1])
        m.update_pars(mod_type="frac_inf_diff", t_half=25.0, bbar=bbar, Phi_im=phi, f_tadv_ext=z[k + "/f_tadv_ext"])
        g = m.gen_g_tp()
        assert rel_err(g, z[k + "/g"]) < 1e-10, k
        assert abs(m.FM_mu - float(z[k + "/FM_mu"])) < 1e-10 * m.FM_mu
        assert abs(m.convolve() - float(z[k + "/C"])) < 1e-10 * abs(float(z[k + "/C"]))
    with pytest.raises(ValueError):
        m.update_pars(mod_type="frac_inf_diff", bbar=1e-3, Phi_im=0.02, f_tadv_ext=np.ones(7))
        m.gen_g_tp()
