"""Config 2 against the posterior traces the reference ships (age_ens_runs_mcmc/conv_traces/*.netcdf: pymc3 3.11.2,
DEMetropolisZ(tune_interval=1000), 3 chains x (10,000 tune + 10,000 draws), run_age_mcmc_utils.py:412-425).

tests/golden/age_traces.json holds summaries of all 22 traces (oracle/gen_trace_fixtures.py, read with the repo's own
NetCDF-4 reader).  What can be pinned: the three single-tracer H3 inversions -- their observation vector is stored in the
trace (`observed_data/like` = ens.mean()) and their error is reconstructible, obs_err = ens.std() + 0.05 obs_mu with the H3
ensemble drawn as N(h3_obs, 0.08 h3_obs) (age_modeling_mcmc.prep.py:96,425; h3_obs = map_dict.pk['H3']).  The CFC / SF6 / He4
ensembles carry the spread of the missing ens_dict.pk, so the joint inversions stay unpinned."""
import json
import os

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FX = os.path.join(ROOT, "tests", "golden", "age_traces.json")
REF = "/root/reference/age_ens_runs_mcmc/conv_traces"
H3_OBS = {"PLM1": 4.868764801408059, "PLM6": 4.154889220496299, "PLM7": 4.323133480432271}     # map_dict.pk['H3']
H3_ENS_ERR, H3_PERR = 0.08, 0.05            # age_modeling_mcmc.prep.py:96 ; run_age_mcmc.py:100-112
TAU_HI_SINGLE = 5000.0                      # single-tracer runs: draws reach 4,999.87 (PLM6.SF6) -> tau1 ~ U(1, 5000)


def fixture():
    return json.load(open(FX))


def h3_sigma(well, obs_mu):
    return H3_ENS_ERR * H3_OBS[well] + H3_PERR * obs_mu


def exact_h3_posterior(well, obs_mu, sigma, tau_hi=TAU_HI_SINGLE):
    """(grid, cdf, mean) of p(tau1 | obs): U(1, tau_hi) x Beta(2, 0.1)[nu_] x StudentT(obs | model(tau1), sigma, 5 + 25 nu_),
    nu_ integrated out with the substitution nu_ = 1 - u^10 (removes the (1 - x)^-0.9 end-point singularity)."""
    import c_oracle
    import np_oracle as O
    from helpers import load_c_in
    X = load_c_in()["H3"].reshape(-1, 1)
    desc = [dict(series=0, rad_accum=False, lam=float(-np.log(0.5) / 12.34))]
    grid = np.unique(np.concatenate([np.arange(1.0, 300.0, 0.05), np.arange(300.0, tau_hi + 1e-9, 1.0)]))
    mu = c_oracle.forward(X, desc, "exponential", False, grid.reshape(-1, 1), ["tau1"])[:, 0]
    ug, uw = np.polynomial.legendre.leggauss(64)
    ug, uw = 0.5 * (ug + 1.0), 0.5 * uw
    post = np.zeros_like(grid)
    for u, w in zip(ug, uw):
        x = 1.0 - u ** 10
        lp = O.logp_studentt(np.array([obs_mu]), mu.reshape(-1, 1), np.array([sigma]), np.full(len(grid), 5.0 + 25.0 * x))
        post += w * 10.0 * x * np.exp(lp)
    post *= np.gradient(grid)
    return grid, np.cumsum(post) / post.sum(), float((grid * post).sum() / post.sum())


def cdf_gap(fx, v, grid, cdf):
    """F_exact(q_ref[p]) - p on the fixture's quantile grid, and the 1-sigma Monte-Carlo error of p at the reference's ESS."""
    p = np.array(fx["qgrid"]) / 100.0
    gap = np.interp(np.array(v["q"]), grid, cdf) - p
    return gap, np.sqrt(p * (1.0 - p) / min(v["ess_bulk"], v["ess_tail"]))


def test_fixture_is_complete_and_self_consistent():
    fx = fixture()
    assert len(fx["traces"]) == 22
    for name, t in fx["traces"].items():
        assert t["chains"] == 3 and t["draws"] == 10000 and t["tuning_steps"] == 10000, name
        assert 60.0 < t["sampling_time"] < 500.0, name                     # 67-493 s per inversion
        v = t["vars"]
        assert abs(v["nu"]["mean"] - (5.0 + 25.0 * v["nu_"]["mean"])) < 1e-9, name          # run_age_mcmc_utils.py:291-292
        if "thalf_cfc" in v:
            assert abs(v["thalf_cfc"]["mean"] - (5.0 + 30.0 * v["thalf_cfc_"]["mean"])) < 1e-9, name      # :333-334
        if "f2" in v:
            assert abs(v["f1"]["mean"] + v["f2"]["mean"] - 1.0) < 1e-12, name                # :305
        assert len(t["obs_mu"]) == len(t["tracers"]), name
    # every inversion of one well saw the same observation of a tracer
    assert fx["traces"]["PLM1.H3.exponential.0"]["obs_mu"][0] == fx["traces"]["PLM1.CFC12.SF6.H3.He4_ter.exponential.123"]["obs_mu"][2]


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree only exists in the build container")
def test_netcdf4_reader_reproduces_the_fixture():
    from noblegas_rtd_mcmc_b200 import netcdf4_reader
    fx = fixture()
    for name in ("PLM7.H3.exponential.0", "PLM6.CFC12.SF6.H3.He4_ter.exp_pist_flow-piston.123"):     # compact and dense groups
        tr = netcdf4_reader.read_trace(os.path.join(REF, name + ".netcdf"))
        t = fx["traces"][name]
        assert sorted(k for k in tr["posterior"] if k not in ("chain", "draw")) == sorted(t["vars"])
        for k, v in t["vars"].items():
            a = tr["posterior"][k]
            assert a.shape == (3, 10000) and a.dtype == np.float64
            assert a.mean() == pytest.approx(v["mean"], rel=1e-13)
            assert np.allclose(np.percentile(a, fx["qgrid"]), v["q"], rtol=1e-13)
        assert np.allclose(tr["observed_data"]["like"], t["obs_mu"], rtol=0, atol=0)
        assert tr["attrs"]["posterior"]["inference_library_version"] == "3.11.2"
        assert float(tr["attrs"]["posterior"]["sampling_time"][0]) == t["sampling_time"]
        assert np.array_equal(tr["sample_stats"]["accepted"].mean(), t["accept_rate"])


@pytest.mark.skipif(not os.path.isdir(REF), reason="the reference tree only exists in the build container")
def test_load_trace_reads_the_reference_netcdf_like_its_own_npz(tmp_path):
    """diagnostics.load_trace on a reference trace -> the layout of our own traces; summary() then gives az.summary's columns."""
    from noblegas_rtd_mcmc_b200 import diagnostics as D
    tr = D.load_trace(os.path.join(REF, "PLM1.CFC12.SF6.H3.He4_ter.exp_pist_flow.123.netcdf"))
    assert sorted(tr["posterior"]) == ["J", "eta1", "lamsf6", "nu", "nu_", "tau1", "thalf_cfc", "thalf_cfc_"]
    assert tr["posterior"]["tau1"].shape == (3, 10000) and tr["observed_data"]["like"].shape == (4,)
    assert tr["attrs"]["inference_library"] == "pymc3" and 296.0 < tr["attrs"]["sampling_time"] < 324.0
    summ = D.summary({k: tr["posterior"][k] for k in ("tau1", "eta1")})
    fxv = fixture()["traces"]["PLM1.CFC12.SF6.H3.He4_ter.exp_pist_flow.123"]["vars"]
    assert summ["tau1"]["ess_bulk"] == pytest.approx(fxv["tau1"]["ess_bulk"]) and summ["eta1"]["r_hat"] == pytest.approx(fxv["eta1"]["r_hat"])
    D.save_trace(str(tmp_path / "t.npz"), tr["posterior"], tr["sample_stats"], tr["attrs"])
    back = D.load_trace(str(tmp_path / "t.npz"))
    assert np.array_equal(back["posterior"]["tau1"], tr["posterior"]["tau1"])


@pytest.mark.parametrize("well", ["PLM1", "PLM6", "PLM7"])
def test_reference_h3_posteriors_match_the_restated_model(well):
    """pymc3's own draws vs the exact posterior of the restated prior x Student-T likelihood x oracle forward model:
    the CDF at the reference's quantiles within 4.5 sigma of the reference's Monte-Carlo error, the mean within 4 MCSE;
    and the check has teeth: dropping the ensemble spread from obs_err moves the CDF by > 0.1."""
    fx = fixture()
    t = fx["traces"]["%s.H3.exponential.0" % well]
    v = t["vars"]["tau1"]
    obs = t["obs_mu"][0]
    grid, cdf, mean = exact_h3_posterior(well, obs, h3_sigma(well, obs))
    gap, sig = cdf_gap(fx, v, grid, cdf)
    assert np.all(np.abs(gap) < 4.5 * sig + 0.003), (well, np.round(gap, 4), np.round(sig, 4))
    assert abs(mean - v["mean"]) < 4.0 * v["mcse_mean"], (well, mean, v["mean"], v["mcse_mean"])
    grid2, cdf2, _ = exact_h3_posterior(well, obs, H3_PERR * obs)
    gap2, _ = cdf_gap(fx, v, grid2, cdf2)
    assert np.abs(gap2).max() > 0.1, (well, np.abs(gap2).max())


@pytest.mark.gpu
@pytest.mark.parametrize("well", ["PLM1", "PLM6", "PLM7"])
def test_device_sampler_reproduces_reference_h3_traces(well):
    """The device sampler with the reference's settings (DEMetropolisZ, tune_interval 1000, tune 10,000 + 10,000 draws; 256
    chains instead of 3) on the real 25,256-lag H3 series against the quantiles of the reference's pymc3 trace."""
    from helpers import real_plan
    from noblegas_rtd_mcmc_b200.sampler import Sampler, prior
    fx = fixture()
    t = fx["traces"]["%s.H3.exponential.0" % well]
    v = t["vars"]["tau1"]
    obs = t["obs_mu"][0]
    plan, _ = real_plan("exponential", False, ["tau1"], ["H3"])
    smp = Sampler([prior("uniform", "tau1", 1.0, TAU_HI_SINGLE), prior("beta", "nu_", 2.0, 0.1)], np.array([obs]),
                  np.array([h3_sigma(well, obs)]), 256, plan=plan, lik="studentt", nu_range=(5.0, 30.0), tune_interval=1000,
                  hist_cap=20000, seed=123423)
    tr = smp.sample(10000, 10000, thin=5).cpu().numpy()
    tau, nu_ = tr[:, :, 0].ravel(), tr[:, :, 1].ravel()
    p = np.array(fx["qgrid"]) / 100.0
    F = np.array([(tau <= q).mean() for q in v["q"]])
    sig = np.sqrt(p * (1.0 - p) / min(v["ess_bulk"], v["ess_tail"]))
    assert np.all(np.abs(F - p) < 4.5 * sig + 0.005), (well, np.round(F - p, 4))
    assert abs(tau.mean() - v["mean"]) < 4.0 * v["mcse_mean"] + 0.01 * v["sd"], (well, tau.mean(), v["mean"])
    vn = t["vars"]["nu"]
    assert abs((5.0 + 25.0 * nu_.mean()) - vn["mean"]) < 4.0 * vn["mcse_mean"] + 0.1, (well, 5.0 + 25.0 * nu_.mean(), vn["mean"])
    # and against the exact posterior (independent of the reference's Monte-Carlo error)
    grid, cdf, mean = exact_h3_posterior(well, obs, h3_sigma(well, obs))
    assert abs(tau.mean() - mean) < 0.03 * v["sd"], (well, tau.mean(), mean)
    qs = np.percentile(tau, [5, 25, 50, 75, 95])
    assert np.abs(np.interp(qs, grid, cdf) - np.array([0.05, 0.25, 0.5, 0.75, 0.95])).max() < 0.015, well


# ------------------------------------------------------------------------------------------------------------------------
# CFC12 / SF6 / He4_ter: obs_err contains the spread of the missing ens_dict.pk.  oracle/fit_obs_err.py estimates ONE scalar
# per (well, tracer) from the single-tracer traces (tests/golden/age_obs_err.json); the joint `.123` inversions of the same well
# use the same errors, so they are out-of-sample: nothing is fitted to them.
OBS_ERR = os.path.join(ROOT, "tests", "golden", "age_obs_err.json")
JOINT_TRACERS = ["CFC12", "SF6", "H3", "He4_ter"]


@pytest.mark.parametrize("tracer", ["CFC12", "SF6", "He4_ter"])
def test_single_tracer_traces_are_consistent_with_one_error_scalar(tracer):
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fit_obs_err as F
    fx, rel = fixture(), json.load(open(OBS_ERR))["rel"]
    grid = F.tau_grid(F.TAU_HI[tracer])
    mu = F.forward_on_grid(tracer, grid)
    for well in ("PLM1", "PLM6", "PLM7"):
        t = fx["traces"]["%s.%s.exponential.0" % (well, tracer)]
        obs, v = t["obs_mu"][0], t["vars"]["tau1"]
        cdf, mean = F.exact_cdf(grid, mu, obs, rel[tracer][well] * obs)
        gap, sig = cdf_gap(fx, v, grid, cdf)
        assert np.all(np.abs(gap) < 4.5 * sig + 0.003), (tracer, well, np.round(gap, 4))
        assert abs(mean - v["mean"]) < 4.0 * v["mcse_mean"] + 0.005 * v["sd"], (tracer, well, mean, v["mean"])


def _joint_run(model, well, nchains=512):
    from helpers import real_plan
    from noblegas_rtd_mcmc_b200 import noble_gas_utils as ng_utils
    from noblegas_rtd_mcmc_b200.sampler import Sampler, prior
    fx, rel = fixture(), json.load(open(OBS_ERR))["rel"]
    t = fx["traces"]["%s.CFC12.SF6.H3.He4_ter.%s.123" % (well, model)]
    obs = np.array(t["obs_mu"])
    sd = np.array([rel[tr][well] for tr in JOINT_TRACERS]) * obs
    epm = model == "exp_pist_flow"
    pn = ["tau1"] + (["eta1"] if epm else []) + ["J", "thalf_cfc", "lamsf6"]
    plan, _ = real_plan(model, False, pn, JOINT_TRACERS)
    J_mu = np.log10(ng_utils.J_flux(Del=1., rho_r=2700, rho_w=1000, U=3.7, Th=10.2, phi=0.05))      # run_age_mcmc.py:182
    # tau1 ~ U(1, 500) in the joint exponential runs (draws reach 499.99); the exp_pist_flow posteriors end far below any bound
    pri = [prior("uniform", "tau1", 1.0, 1000.0 if epm else 500.0), prior("beta", "nu_", 2.0, 0.1), prior("normal", "J", J_mu, 0.33)]
    if epm:
        pri.append(prior("uniform", "eta1", 1.0, 5.0))
    pri += [prior("beta", "thalf_cfc", 2.0, 2.0, lo=5.0, hi=35.0), prior("halfnormal", "lamsf6", 0.5 / 3)]
    smp = Sampler(pri, obs, sd, nchains, plan=plan, lik="studentt", nu_range=(5.0, 30.0), tune_interval=1000, hist_cap=20000,
                  seed=123423)
    tr = smp.sample(10000, 10000, thin=5).cpu().numpy()           # [draw, chain, dim]
    smp.close()
    return fx, t, [q["target"] for q in pri], tr


def _compare(fx, t, names, tr, skip=(), extra=0.01, nsig=4.5):
    p = np.array(fx["qgrid"]) / 100.0
    worst = {}
    for i, nm in enumerate(names):
        if nm in skip:
            continue
        v = t["vars"][nm]
        a = tr[:, :, i].ravel()
        F = np.array([(a <= q).mean() for q in v["q"]])
        sig = np.sqrt(p * (1.0 - p) / min(v["ess_bulk"], v["ess_tail"]))
        worst[nm] = float((np.abs(F - p) - nsig * sig).max())
        assert np.all(np.abs(F - p) < nsig * sig + extra), (nm, np.round(F - p, 4), np.round(sig, 4))
    return worst


@pytest.mark.gpu
@pytest.mark.parametrize("well", ["PLM1", "PLM6", "PLM7"])
def test_joint_exponential_inversion_matches_reference_trace(well):
    """`<well>.CFC12.SF6.H3.He4_ter.exponential.123`: tau1, J, thalf_cfc, lamsf6 of the device sampler against the quantiles of
    the reference's pymc3 trace, out of sample (errors from the single-tracer traces).  nu_ piles up at 1 and reacts to the
    fitted scalars, so only its mean is bounded."""
    fx, t, names, tr = _joint_run("exponential", well)
    _compare(fx, t, names, tr, skip=("nu_",))
    nu_ = tr[:, :, names.index("nu_")].mean()
    assert abs(nu_ - t["vars"]["nu_"]["mean"]) < 0.15, (nu_, t["vars"]["nu_"]["mean"])


@pytest.mark.gpu
@pytest.mark.parametrize("well", ["PLM6", "PLM7", "PLM1"])
def test_joint_exp_pist_flow_inversion_matches_reference_trace(well):
    """BASELINE config 2 itself (`<well>.CFC12.SF6.H3.He4_ter.exp_pist_flow.123`).  The PLM1 posterior is bimodal in
    (tau1, eta1); DE-MC-Z chains stay in the mode they tune into, and the reference's three chains all sit in the
    eta1 ~ 1.2 mode -- the comparison is made with the chains of ours that sit in the same mode."""
    fx, t, names, tr = _joint_run("exp_pist_flow", well)
    ie = names.index("eta1")
    ref_eta = t["vars"]["eta1"]
    in_mode = np.abs(tr[:, :, ie].mean(axis=0) - ref_eta["mean"]) < 4.0 * ref_eta["sd"] + 0.1
    assert in_mode.sum() >= 64, in_mode.sum()
    if well != "PLM1":
        assert in_mode.mean() > 0.95, in_mode.mean()
    _compare(fx, t, names, tr[:, in_mode, :], skip=("nu_",))


@pytest.mark.gpu
@pytest.mark.parametrize("tracer", ["CFC12", "SF6", "He4_ter"])
def test_device_sampler_reproduces_single_tracer_traces(tracer):
    """The nine remaining `<well>.<tracer>.exponential.0` inversions on the device sampler (reference settings, 256 chains) against
    the quantiles of the pymc3 traces; obs_err is the one scalar per trace of tests/golden/age_obs_err.json (in sample)."""
    import sys
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import fit_obs_err as F
    from helpers import real_plan
    from noblegas_rtd_mcmc_b200.sampler import Sampler, prior
    fx, rel = fixture(), json.load(open(OBS_ERR))["rel"]
    p = np.array(fx["qgrid"]) / 100.0
    plan, _ = real_plan("exponential", False, ["tau1"], [tracer])
    for well in ("PLM1", "PLM6", "PLM7"):
        t = fx["traces"]["%s.%s.exponential.0" % (well, tracer)]
        obs, v = t["obs_mu"][0], t["vars"]["tau1"]
        smp = Sampler([prior("uniform", "tau1", 1.0, F.TAU_HI[tracer]), prior("beta", "nu_", 2.0, 0.1)], np.array([obs]),
                      np.array([rel[tracer][well] * obs]), 256, plan=plan, lik="studentt", nu_range=(5.0, 30.0),
                      tune_interval=1000, hist_cap=20000, seed=123423)
        tau = smp.sample(10000, 10000, thin=5).cpu().numpy()[:, :, 0].ravel()
        smp.close()
        Fq = np.array([(tau <= q).mean() for q in v["q"]])
        sig = np.sqrt(p * (1.0 - p) / min(v["ess_bulk"], v["ess_tail"]))
        assert np.all(np.abs(Fq - p) < 4.5 * sig + 0.005), (tracer, well, np.round(Fq - p, 4))
        assert abs(tau.mean() - v["mean"]) < 4.0 * v["mcse_mean"] + 0.01 * v["sd"], (tracer, well, tau.mean(), v["mean"])

