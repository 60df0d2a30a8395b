"""Quick device-side timing of the cfg-3 forward+loglik kernel (development aid; bench.py is the contract)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from noblegas_rtd_mcmc_b200 import synthetic
from helpers import synth_plan

B = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
cfgs = {"epm_dm": ("exp_pist_flow", "dispersion", list(synthetic.PAR_NAMES_CFG3)),
        "epm": ("exp_pist_flow", False, ["tau1", "eta1", "J"]),
        "dm": ("dispersion", False, ["tau1", "D1", "J"]),
        "epm_epm": ("exp_pist_flow", "exp_pist_flow", ["tau1", "tau2", "f1", "f2", "eta1", "eta2", "J"]),
        "dm_dm": ("dispersion", "dispersion", ["tau1", "tau2", "f1", "f2", "D1", "D2", "J"])}
th7 = synthetic.theta_cfg3(B, 0)
cols = dict(zip(synthetic.PAR_NAMES_CFG3, th7.T))
cols["eta2"] = cols["eta1"]; cols["D1"] = cols["D2"]
for name, (m1, m2, pn) in cfgs.items():
    plan, _, _ = synth_plan(m1, m2, pn)
    theta = torch.from_numpy(np.ascontiguousarray(np.stack([cols[p] for p in pn], 1))).cuda()
    obs = np.ones(7); sd = np.ones(7) * 0.05
    logp = torch.empty(B, dtype=torch.float64, device="cuda")
    for _ in range(3):
        plan.forward_loglik_dev(theta, pn, obs, sd, "normal", logp_t=logp)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 20
    e0.record()
    for _ in range(n):
        plan.forward_loglik_dev(theta, pn, obs, sd, "normal", logp_t=logp)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    ncomp = 2 if m2 else 1
    flops = 2.0 * 840 * 8 * ncomp * B   # 8 columns incl. normalisation
    print("%-8s B=%d  %.3f ms/launch  %.3e tracer-evals/s (7 tracers)  %.2f TFLOP/s algorithmic (8 cols)  nan=%d" % (
        name, B, ms, B * 7 / ms * 1e3, flops / ms / 1e9, int(torch.isnan(logp).sum())))

# fused Metropolis steps in the persistent sampler kernel (cfg 3, 7 sampler dims, Student-T, DE-MC-Z)
from noblegas_rtd_mcmc_b200.sampler import Sampler, prior
pn = list(synthetic.PAR_NAMES_CFG3)
plan, _, _ = synth_plan("exp_pist_flow", "dispersion", pn)
truth = np.array([[180.0, 1500.0, 0.6, 0.4, 1.8, 0.4, synthetic.LOG10_J_MONTHLY]])
obs = plan.forward_host(truth, pn)[0]; sd = 0.05 * np.abs(obs)
pri = [prior("uniform", "tau1", 12, 12000), prior("beta", "nu_", 2.0, 0.1), prior("normal", "J", synthetic.LOG10_J_MONTHLY, 0.33),
       prior("uniform", "tau2", 600, 180000), prior("uniform", "f1", 0.01, 0.99), prior("uniform", "eta1", 1, 5),
       prior("uniform", "D2", 0.01, 2.0)]
for lik in ("normal", "studentt"):
    smp = Sampler(pri, obs, sd, B, plan=plan, lik=lik, nu_range=(5.0, 30.0), f2_from_f1=True, tune_interval=100, hist_cap=256, seed=1)
    smp.run(20, tune=True); torch.cuda.synchronize()
    for K in (50, 200):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); smp.run(K, tune=True); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print("sampler %-8s K=%3d  %.4f ms/step  %.3e evals/s (6 counted)  %.2f TF (7 col)  acc %.3f" % (
            lik, K, ms, B * 6 / ms * 1e3, 2.0 * 840 * 7 * 2 * B / ms / 1e9, float(smp.get("accepted").mean()) / smp.info()["step"]))
    smp.close()
