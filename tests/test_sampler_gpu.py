"""GPU: the device-resident sampler against (a) the numpy restatement driven by the same Philox stream (trajectory
parity, step by step), (b) shard invariance, (c) the reference's shipped posterior summaries for config 1
(ng_interp/ng_optPLM*.csv: the only posterior known answers in the reference) and (d) a posterior computed by brute-force
quadrature with the oracle forward model."""
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_device_philox_known_answers():
    from noblegas_rtd_mcmc_b200.sampler import philox4x32_10
    assert list(philox4x32_10([0] * 4, [0] * 2)) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert list(philox4x32_10([0xffffffff] * 4, [0xffffffff] * 2)) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert list(philox4x32_10([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0])) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def _ng_setup():
    from helpers import GOLD
    from noblegas_rtd_mcmc_b200.noble_gas_mcmc import mcmc_model
    fx = json.load(open(os.path.join(GOLD, "ng_posterior.json")))
    return fx, mcmc_model


def _natural(priors, Q):
    import np_sampler as S
    return np.array([[S.transform_dim(pr, q[d])[0] for d, pr in enumerate(priors)] for q in Q])


def test_ng_trajectory_matches_numpy_restatement():
    """4 chains x 360 steps of DE-MC-Z on the closed-equilibrium model, two lambda-tuning points and stop_tuning
    inside the window: device trace == numpy restatement with the same Philox counters."""
    import np_oracle as O
    import np_sampler as S
    from noblegas_rtd_mcmc_b200.sampler import Sampler
    fx, mcmc_model = _ng_setup()
    mdl = mcmc_model(fx["wells"]["PLM1"]["obs"], mcmc_model.well_elev["PLM1"])
    pri = mdl.build_priors()
    smp = Sampler(pri, mdl.obs_mu, mdl.obs_sd, 4, plan=None, gases=mdl.gases, lik="studentt", nu_range=(1.0, 30.0),
                  tune_interval=100, hist_cap=1000, seed=99)
    tr = [smp.run(240, tune=True, record=True, keep_trace=True).cpu().numpy()]
    smp.stop_tuning()
    tr.append(smp.run(120, tune=False, record=True, keep_trace=True).cpu().numpy())
    dev = np.concatenate(tr, axis=0)                                    # [step, chain, dim]

    def logp_model(v):
        mu = O.ce_exc(mdl.gases, v["E"], (v["E"] - v["b"]) / v["m"], 10 ** v["log10Ae"], 10 ** v["log10F"], True)[0]
        return S.studentt_logp(mdl.obs_mu, mu, mdl.obs_sd, 1.0 + 29.0 * v["nu_"])
    for c in range(4):
        Q, LP, AC = S.run_chain(pri, logp_model, 360, seed=99, chain=c, tune_steps=240, tune_interval=100)
        want = _natural(pri, Q)
        assert np.allclose(dev[:, c, :], want, rtol=1e-8, atol=0), "chain %d diverges at step %d" % (
            c, int(np.argmax(~np.isclose(dev[:, c, :], want, rtol=1e-8).all(axis=1))))
        assert 0.02 < AC.mean() < 0.98


def test_age_trajectory_matches_numpy_restatement():
    """EPM + dispersion mixture, 7 tracers, Student-T likelihood: 3 chains x 150 steps through the fused kernel."""
    import np_oracle as O
    import np_sampler as S
    from helpers import synth_plan
    from noblegas_rtd_mcmc_b200 import synthetic
    from noblegas_rtd_mcmc_b200.sampler import Sampler, prior
    pn = list(synthetic.PAR_NAMES_CFG3)
    plan, series, tab = synth_plan("exp_pist_flow", "dispersion", pn)
    truth = np.array([[180.0, 1500.0, 0.6, 0.4, 1.8, 0.4, synthetic.LOG10_J_MONTHLY]])
    obs = plan.forward_host(truth, pn)[0]
    sd = 0.05 * np.abs(obs)
    pri = [prior("uniform", "tau1", 12, 12000), prior("beta", "nu_", 2.0, 0.1), prior("normal", "J", synthetic.LOG10_J_MONTHLY, 0.33),
           prior("uniform", "tau2", 600, 180000), prior("uniform", "f1", 0.01, 0.99), prior("uniform", "eta1", 1, 5),
           prior("uniform", "D2", 0.01, 2.0)]
    q0 = [-3.0, 2.0, synthetic.LOG10_J_MONTHLY, -4.5, 0.3, -1.0, -1.2]      # start inside the informative region
    smp = Sampler(pri, obs, sd, 3, plan=plan, lik="studentt", nu_range=(5.0, 30.0), f2_from_f1=True, tune_interval=50,
                  hist_cap=500, seed=2024, q0=q0, scaling=0.01)
    dev = smp.run(150, tune=True, record=True, keep_trace=True).cpu().numpy()

    def logp_model(v):
        theta = np.array([[v["tau1"], v["tau2"], v["f1"], 1.0 - v["f1"], v["eta1"], v["D2"], v["J"]]])
        mu = []
        for t in synthetic.TRACERS_CFG3:
            d = tab[t]
            s = series[d["series"]] if d["series"] is not None else np.zeros(840)
            mu.append(O.forward_mod(theta, pn, t, s, "exp_pist_flow", "dispersion", t_half=d.get("t_half", False),
                                    rad_accum=d.get("rad_accum", False))[0])
        return S.studentt_logp(obs, np.array(mu), sd, 5.0 + 25.0 * v["nu_"])
    nacc = 0
    for c in range(3):
        Q, LP, AC = S.run_chain(pri, logp_model, 150, seed=2024, chain=c, tune_steps=150, tune_interval=50, scaling=0.01, q0=q0)
        want = _natural(pri, Q)
        assert np.allclose(dev[:, c, :], want, rtol=1e-7, atol=0), "chain %d diverges at step %d" % (
            c, int(np.argmax(~np.isclose(dev[:, c, :], want, rtol=1e-7).all(axis=1))))
        nacc += AC.sum()
    assert nacc > 20                                  # the window exercises accepted moves, not only rejections


def test_age_sampler_table_variants_agree():
    """k_mcmc_age has two instantiations per dispersion plan: 2,048-entry exp table + compact per-chain records (samplers of
    up to 8 dimensions whose layout fits next to the resident lag tables: the default for this plan) and 128-entry table +
    ND_MAX records (NGRTD_MCMC_TB11=0 forces it).  Their weights differ by ~1e-13, so 96 chains x 120 steps from the same
    Philox streams must take the same accept decisions and end in the same states."""
    import subprocess
    import sys
    probe = os.path.join(os.path.dirname(os.path.abspath(__file__)), "sampler_variant_probe.py")
    res = []
    for flag in ("1", "0"):
        env = dict(os.environ, NGRTD_MCMC_TB11=flag)
        out = subprocess.run([sys.executable, probe], env=env, capture_output=True, text=True, timeout=600)
        assert out.returncode == 0, out.stderr[-2000:]
        res.append(json.loads(out.stdout.strip().splitlines()[-1]))
    a, b = res
    assert np.array_equal(np.array(a["accepted"]), np.array(b["accepted"]))
    assert np.sum(a["accepted"]) > 100
    assert np.allclose(a["q"], b["q"], rtol=1e-9, atol=0)
    assert np.allclose(a["logp"], b["logp"], rtol=1e-9, atol=1e-9)


def test_shard_invariance_bitwise():
    """A chain's trajectory is a pure function of (seed, global chain id): 64 chains in one sampler == two shards."""
    import torch
    from noblegas_rtd_mcmc_b200.sampler import Sampler
    fx, mcmc_model = _ng_setup()
    mdl = mcmc_model(fx["wells"]["PLM7"]["obs"], mcmc_model.well_elev["PLM7"])
    kw = dict(plan=None, gases=mdl.gases, lik="studentt", nu_range=(1.0, 30.0), tune_interval=50, hist_cap=256, seed=5)
    a = Sampler(mdl.build_priors(), mdl.obs_mu, mdl.obs_sd, 64, **kw)
    b0 = Sampler(mdl.build_priors(), mdl.obs_mu, mdl.obs_sd, 24, chain_offset=0, **kw)
    b1 = Sampler(mdl.build_priors(), mdl.obs_mu, mdl.obs_sd, 40, chain_offset=24, **kw)
    for s in (a, b0, b1):
        s.run(300, tune=True)
        s.stop_tuning()
        s.run(100, tune=False, record=True)
    qa = a.get("q")
    qb = torch.cat([b0.get("q"), b1.get("q")], dim=0)
    assert torch.equal(qa, qb)
    assert torch.equal(a.get("mean"), torch.cat([b0.get("mean"), b1.get("mean")], dim=0))
    assert a.info() == dict(step=400, ndraws=100, hist_start=270)


@pytest.mark.parametrize("well", ["PLM1", "PLM7", "PLM6"])
def test_ng_posterior_matches_reference_summary(well):
    """Config 1 with the reference's sampler settings (DEMetropolisZ, tune 10,000, tune_interval 5,000, 50,000 draws;
    ng_interp/noble_gas_mcmc.py:408-415) on 256 chains instead of 4, so that OUR Monte-Carlo error is small.
    Posterior mean / median / sd / 94 % HDI must agree with ng_optPLM*.csv within the combined Monte-Carlo error
    (the reference's own MCSE dominates: its bulk ESS is only 954-7,239)."""
    from noblegas_rtd_mcmc_b200 import diagnostics as D
    fx, mcmc_model = _ng_setup()
    w = fx["wells"][well]
    mdl = mcmc_model(w["obs"], mcmc_model.well_elev[well])
    res = mdl.sample(chains=256, tune=10000, draws=50000, tune_interval=5000, random_seed=123423, thin=10)
    post = res["posterior"]
    acc = res["sample_stats"]["accept_rate"]
    assert 0.05 < acc.mean() < 0.6
    for var in ("T", "E", "Ae", "F", "m", "b", "nu"):
        ref = w["summary"][var]
        a = post[var]
        ours_mcse = a.std() / np.sqrt(max(D.ess_mean(a), 10.0))
        tol = 5.0 * np.hypot(ref["mcse_mean"], ours_mcse) + 0.02 * ref["sd"]
        assert abs(a.mean() - ref["mean"]) < tol, (well, var, "mean", a.mean(), ref["mean"], tol)
        assert abs(np.median(a) - ref["median"]) < 1.5 * tol, (well, var, "median", np.median(a), ref["median"])
        assert abs(a.std() - ref["sd"]) < 0.15 * ref["sd"] + 5 * ref["mcse_sd"], (well, var, "sd", a.std(), ref["sd"])
        lo, hi = D.hdi(a)
        assert abs(lo - ref["hdi_3%"]) < 0.2 * ref["sd"] and abs(hi - ref["hdi_97%"]) < 0.2 * ref["sd"], (well, var, lo, hi)
    assert D.rhat(post["T"][:64]) < 1.02


def test_age_posterior_matches_quadrature():
    """One-tracer exponential model (the `.0` configuration of the reference): posterior of tau1 from the sampler vs
    brute-force quadrature of prior x Gaussian likelihood with the oracle forward model on a fine grid."""
    import np_oracle as O
    from helpers import load_c_in, real_plan
    from noblegas_rtd_mcmc_b200 import diagnostics as D
    from noblegas_rtd_mcmc_b200.sampler import Sampler, prior
    pn = ["tau1"]
    plan, C = real_plan("exponential", False, pn, ["CFC12"], L=2000)
    obs, sd = np.array([36.382794]), np.array([36.382794 * 0.05 + 1.5])
    smp = Sampler([prior("uniform", "tau1", 1.0, 1000.0)], obs, sd, 512, plan=plan, lik="normal", tune_interval=500,
                  hist_cap=6000, seed=11)
    tr = smp.sample(3000, 3000, thin=2).cpu().numpy()[:, :, 0]
    grid = np.linspace(1.0, 1000.0, 40001)
    mu = O.forward_mod(grid.reshape(-1, 1), pn, "CFC12", C["CFC12"], "exponential", False)
    w = np.exp(O.logp_normal(obs, mu.reshape(-1, 1), sd) - O.logp_normal(obs, mu.reshape(-1, 1), sd).max())
    w /= w.sum()
    qm = float((grid * w).sum())
    qs = float(np.sqrt(((grid - qm) ** 2 * w).sum()))
    a = tr.T
    mcse = a.std() / np.sqrt(D.ess_mean(a))
    assert abs(a.mean() - qm) < 5 * mcse + 0.01 * qs, (a.mean(), qm, mcse)
    assert abs(a.std() - qs) < 0.08 * qs, (a.std(), qs)


def test_observation_groups_config4():
    """Config 4 (wells x ensemble members): one sampler with per-group observation rows == independent samplers per
    group (bit-identical trajectories, because chains are keyed by global id), for both kernels."""
    import torch
    from helpers import synth_plan
    from noblegas_rtd_mcmc_b200 import synthetic
    from noblegas_rtd_mcmc_b200.sampler import Sampler, prior
    fx, mcmc_model = _ng_setup()
    wells = ["PLM1", "PLM6", "PLM7"]
    mdls = [mcmc_model(fx["wells"][w]["obs"], mcmc_model.well_elev["PLM1"]) for w in wells]
    kw = dict(plan=None, gases=mdls[0].gases, lik="studentt", nu_range=(1.0, 30.0), tune_interval=50, hist_cap=400, seed=3)
    joint = Sampler(mdls[0].build_priors(), mdls[0].obs_mu, mdls[0].obs_sd, 3 * 40, **kw)
    joint.set_obs_groups(np.stack([m.obs_mu for m in mdls]), np.stack([m.obs_sd for m in mdls]), 40)
    joint.run(200, tune=True)
    qj = joint.get("q")
    for g, m in enumerate(mdls):
        solo = Sampler(m.build_priors(), m.obs_mu, m.obs_sd, 40, chain_offset=40 * g, **kw)
        solo.run(200, tune=True)
        assert torch.equal(qj[40 * g:40 * (g + 1)], solo.get("q")), wells[g]
    # age model, 2 groups x 24 chains through the fused kernel (groups straddle the 16-chain units)
    pn = list(synthetic.PAR_NAMES_CFG3)
    plan, _, _ = synth_plan("exp_pist_flow", "dispersion", pn)
    truths = np.array([[180.0, 1500.0, 0.6, 0.4, 1.8, 0.4, synthetic.LOG10_J_MONTHLY],
                       [90.0, 2500.0, 0.3, 0.7, 1.2, 0.9, synthetic.LOG10_J_MONTHLY]])
    obs = plan.forward_host(truths, pn)
    sd = 0.05 * np.abs(obs)
    pri = [prior("uniform", "tau1", 12, 12000), prior("beta", "nu_", 2.0, 0.1), prior("normal", "J", synthetic.LOG10_J_MONTHLY, 0.33),
           prior("uniform", "tau2", 600, 180000), prior("uniform", "f1", 0.01, 0.99), prior("uniform", "eta1", 1, 5),
           prior("uniform", "D2", 0.01, 2.0)]
    q0 = [-3.0, 2.0, synthetic.LOG10_J_MONTHLY, -4.5, 0.3, -1.0, -1.2]
    akw = dict(plan=plan, lik="studentt", nu_range=(5.0, 30.0), f2_from_f1=True, tune_interval=25, hist_cap=200, seed=8, q0=q0, scaling=0.01)
    joint = Sampler(pri, obs[0], sd[0], 48, **akw)
    joint.set_obs_groups(obs, sd, 24)
    joint.run(60, tune=True)
    for g in range(2):
        solo = Sampler(pri, obs[g], sd[g], 24, chain_offset=24 * g, **akw)
        solo.run(60, tune=True)
        assert torch.equal(joint.get("q")[24 * g:24 * (g + 1)], solo.get("q")), g
    with pytest.raises(Exception):
        joint.set_obs_groups(obs, sd, 8)          # 48 chains would need 6 groups


def test_checkpoint_resume_is_bit_identical(tmp_path):
    """state_dict -> npz -> a NEW sampler -> continue: same trajectory, statistics and trace as an
    uninterrupted run (both kernels)."""
    import torch
    from helpers import synth_plan
    from noblegas_rtd_mcmc_b200 import synthetic
    from noblegas_rtd_mcmc_b200.sampler import Sampler, prior
    fx, mcmc_model = _ng_setup()
    mdl = mcmc_model(fx["wells"]["PLM1"]["obs"], mcmc_model.well_elev["PLM1"])
    kw = dict(plan=None, gases=mdl.gases, lik="studentt", nu_range=(1.0, 30.0), tune_interval=64, hist_cap=128, seed=21)
    pn = list(synthetic.PAR_NAMES_CFG3)
    plan, _, _ = synth_plan("exp_pist_flow", "dispersion", pn)
    obs = plan.forward_host(np.array([[180.0, 1500.0, 0.6, 0.4, 1.8, 0.4, synthetic.LOG10_J_MONTHLY]]), pn)[0]
    pri = [prior("uniform", "tau1", 12, 12000), prior("beta", "nu_", 2.0, 0.1), prior("normal", "J", synthetic.LOG10_J_MONTHLY, 0.33),
           prior("uniform", "tau2", 600, 180000), prior("uniform", "f1", 0.01, 0.99), prior("uniform", "eta1", 1, 5),
           prior("uniform", "D2", 0.01, 2.0)]
    akw = dict(plan=plan, lik="studentt", nu_range=(5.0, 30.0), f2_from_f1=True, tune_interval=20, hist_cap=48, seed=4, scaling=0.01,
               q0=[-3.0, 2.0, synthetic.LOG10_J_MONTHLY, -4.5, 0.3, -1.0, -1.2])
    for make, n1, n2 in ((lambda: Sampler(mdl.build_priors(), mdl.obs_mu, mdl.obs_sd, 33, **kw), 150, 170),
                         (lambda: Sampler(pri, obs, 0.05 * np.abs(obs), 19, **akw), 50, 45)):
        a = make()
        a.run(n1, tune=True)
        a.stop_tuning()
        a.run(n2, tune=False, record=True)
        ta = a.run(30, tune=False, record=True, keep_trace=True)
        b = make()
        b.run(n1, tune=True)
        b.stop_tuning()
        b.run(n2 - 20, tune=False, record=True)
        np.savez(tmp_path / "ckpt.npz", **b.state_dict())
        c = make()
        c.load_state_dict(dict(np.load(tmp_path / "ckpt.npz")))
        assert c.info() == b.info()
        c.run(20, tune=False, record=True)
        tc = c.run(30, tune=False, record=True, keep_trace=True)
        assert torch.equal(ta, tc)
        for k in ("q", "logp", "lamb", "mean", "m2", "accepted"):
            assert torch.equal(a.get(k), c.get(k)), k
