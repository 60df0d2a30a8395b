"""Probe of the one converged binary-mixing trace of the reference (PLM1 ... exponential-piston.123) against the device sampler under
varied observation errors (needs a GPU); development aid, see DESIGN.md section 2."""
import sys, os, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, ROOT + "/tests"); sys.path.insert(0, ROOT + "/oracle")
import numpy as np
import test_reference_traces as T
from helpers import real_plan
from noblegas_rtd_mcmc_b200 import noble_gas_utils as ng_utils
from noblegas_rtd_mcmc_b200.sampler import Sampler, prior
fx, rel = T.fixture(), json.load(open(T.OBS_ERR))["rel"]
t = fx["traces"]["PLM1.CFC12.SF6.H3.He4_ter.exponential-piston.123"]
obs = np.array(t["obs_mu"]); sd = np.array([rel[tr]["PLM1"] for tr in T.JOINT_TRACERS]) * obs
pn = ["tau1", "tau2", "f1", "f2", "J", "thalf_cfc", "lamsf6"]
plan, _ = real_plan("exponential", "piston", pn, T.JOINT_TRACERS)
J_mu = np.log10(ng_utils.J_flux(Del=1., rho_r=2700, rho_w=1000, U=3.7, Th=10.2, phi=0.05))
pri = [prior("uniform", "tau1", 1.0, 1000.0), prior("beta", "nu_", 2.0, 0.1), prior("normal", "J", J_mu, 0.33),
       prior("uniform", "tau2", 50.0, 15000.0), prior("uniform", "f1", 0.01, 0.99),
       prior("beta", "thalf_cfc", 2.0, 2.0, lo=5.0, hi=35.0), prior("halfnormal", "lamsf6", 0.5 / 3)]
p = np.array(fx["qgrid"]) / 100
smp = Sampler(pri, obs, sd, 1024, plan=plan, lik="studentt", nu_range=(5.0, 30.0), f2_from_f1=True, tune_interval=1000, hist_cap=20000, seed=123423)
tr = smp.sample(10000, 10000, thin=5).cpu().numpy(); smp.close()
names = [q["target"] for q in pri]
t2 = tr[:, :, names.index("tau2")]
young = (t2 < 200.0).mean(axis=0)                 # per chain: share of draws with the piston component younger than 200 yr
print("share of draws with tau2 < 200 per chain: quantiles", np.round(np.percentile(young, [0, 10, 25, 50, 75, 90, 100]), 3))
for label, sel in (("all chains", young >= 0), ("chains with < 2 %% young-piston draws (%d)" % (young < 0.02).sum(), young < 0.02),
                   ("chains with > 20 %% young-piston draws (%d)" % (young > 0.2).sum(), young > 0.2)):
    print(label)
    if sel.sum() == 0:
        continue
    for i, nm in enumerate(names):
        v = t["vars"][nm]; a = tr[:, sel, i].ravel()
        F = np.array([(a <= x).mean() for x in v["q"]])
        print("     %-10s ours %10.4g +- %-9.3g ref %10.4g +- %-9.3g  max|F-p| %.3f" % (nm, a.mean(), a.std(), v["mean"], v["sd"], np.abs(F - p).max()))
for label, sdv in (("as fitted", sd), ("SF6 error x 100 (uninformative)", sd * np.array([1, 100.0, 1, 1])), ("CFC12 error x 100", sd * np.array([100.0, 1, 1, 1])),
                   ("H3 error x 100", sd * np.array([1, 1, 100.0, 1])), ("He4 error x 100", sd * np.array([1, 1, 1, 100.0]))):
    smp = Sampler(pri, obs, sdv, 512, plan=plan, lik="studentt", nu_range=(5.0, 30.0), f2_from_f1=True, tune_interval=1000, hist_cap=20000, seed=123423)
    tr = smp.sample(10000, 10000, thin=5).cpu().numpy(); smp.close()
    print(label)
    for i, q in enumerate(pri):
        nm = q["target"]; v = t["vars"][nm]; a = tr[:, :, i].ravel()
        F = np.array([(a <= x).mean() for x in v["q"]])
        print("     %-10s ours %10.4g +- %-9.3g ref %10.4g +- %-9.3g  max|F-p| %.3f" % (nm, a.mean(), a.std(), v["mean"], v["sd"], np.abs(F - p).max()))
