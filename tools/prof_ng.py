"""A few launches of the noble-gas CE sampler kernel (config 1; target of ncu captures). Development aid."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from noblegas_rtd_mcmc_b200.noble_gas_mcmc import mcmc_model
from noblegas_rtd_mcmc_b200.sampler import Sampler
fx = json.load(open(os.path.join(ROOT, "tests", "golden", "ng_posterior.json")))["wells"]["PLM1"]
mdl = mcmc_model(fx["obs"], mcmc_model.well_elev["PLM1"])
NGC = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
K = int(sys.argv[2]) if len(sys.argv) > 2 else 200
ngs = Sampler(mdl.build_priors(), mdl.obs_mu, mdl.obs_sd, NGC, plan=None, gases=mdl.gases, lik="studentt",
              nu_range=(1.0, 30.0), tune_interval=5000, hist_cap=2048, seed=123423)
ngs.run(2100, tune=True)          # fill the history ring
ngs.stop_tuning()
for _ in range(3):
    ngs.run(K, tune=False, record=True)
torch.cuda.synchronize()
print("ok", float(ngs.get("accepted").float().mean()))
