"""Host side of the device-resident batched sampler (libngrtd.so: ngrtd_sampler_*).

Plays the role pymc3's `DEMetropolisZ` + `mc.sample` play in the reference (age_ens_runs_mcmc/run_age_mcmc_utils.py:
407-429, ng_interp/noble_gas_mcmc.py:402-422) for thousands to millions of independent chains: a launch advances every
chain by `nsteps` whole Metropolis steps on the GPU.  PyTorch only owns the device buffers (trace, statistics).
"""
import ctypes

import numpy as np

from . import _lib


def prior(kind, target, p0, p1=0.0, lo=0.0, hi=1.0):
    """One sampler dimension.  kind: 'uniform'(a,b) | 'beta'(alpha,beta -> [lo,hi]) | 'normal'(mu,sigma) | 'halfnormal'(sigma).
    target: parameter name ('tau1', ..., 'lamsf6', 'nu_') for the age model, or 'log10Ae','log10F','E','m','b','nu_'."""
    return dict(kind=kind, target=target, p0=float(p0), p1=float(p1), lo=float(lo), hi=float(hi))


class Sampler(object):
    """All chains of one shard.  `plan` = _lib.Plan for the age model, None for the noble-gas CE model."""

    def __init__(self, priors, obs_mu, obs_sd, nchains, plan=None, gases=None, lik="studentt", nu_range=None,
                 nu_fixed=30.0, f2_from_f1=False, proposal="uniform", de_mcz=True, tune_target="lambda",
                 tune_interval=1000, scaling=0.001, lamb=0.0, tune_drop_fraction=0.9, hist_cap=20000, seed=123423,
                 chain_offset=0, q0=None, device=-1):
        import torch
        cfg = _lib.SamplerCfg()
        cfg.ndim = len(priors)
        if not 1 <= cfg.ndim <= 10:
            raise ValueError("1..10 sampler dimensions are supported")
        tmap = _lib.NG_TARGET if plan is None else dict({k: v for k, v in _lib.SLOT.items() if k != "f1_f2c"}, nu_=_lib.VAL_NU)
        self.names = []
        nu_sampled = False
        for i, p in enumerate(priors):
            if p["kind"] not in _lib.PRIOR_KIND:
                raise ValueError("unknown prior kind %r" % (p["kind"],))
            if p["target"] not in tmap:
                raise ValueError("unknown prior target %r (known: %s)" % (p["target"], sorted(tmap)))
            cfg.prior[i] = _lib.Prior(_lib.PRIOR_KIND[p["kind"]], tmap[p["target"]], p["p0"], p["p1"], p["lo"], p["hi"])
            self.names.append(p["target"])
            nu_sampled |= p["target"] == "nu_"
        obs_mu = np.asarray(obs_mu, dtype=np.float64).ravel()
        obs_sd = np.asarray(obs_sd, dtype=np.float64).ravel()
        if len(obs_mu) != len(obs_sd) or not 1 <= len(obs_mu) <= 8:
            raise ValueError("obs_mu / obs_sd must have the same length in 1..8")
        cfg.lik_kind = _lib.LIK[lik]
        cfg.nu_sampled = int(nu_sampled)
        if nu_sampled and nu_range is None:
            raise ValueError("nu_ is sampled: nu_range=(lo, hi) is required")
        cfg.nu_lo, cfg.nu_hi = (nu_range if nu_range is not None else (0.0, 0.0))
        cfg.nu_fixed = float(nu_fixed)
        cfg.nobs = len(obs_mu)
        for i in range(len(obs_mu)):
            cfg.obs_mu[i], cfg.obs_sd[i] = obs_mu[i], obs_sd[i]
        cfg.f2_from_f1 = int(bool(f2_from_f1))
        cfg.proposal_dist = {"uniform": 0, "normal": 1}[proposal]
        cfg.de_mcz = int(bool(de_mcz))
        cfg.tune_target = {"lambda": 0, "scaling": 1}[tune_target]
        cfg.tune_interval = int(tune_interval)
        cfg.scaling = float(scaling)
        cfg.lamb = float(lamb)
        cfg.tune_drop_fraction = float(tune_drop_fraction)
        cfg.hist_cap = int(hist_cap)
        cfg.seed = int(seed)
        cfg.chain_offset = int(chain_offset)
        if plan is None:
            if not gases:
                raise ValueError("the noble-gas model needs the list of modelled gases")
            cfg.ngas = len(gases)
            for i, g in enumerate(gases):
                cfg.gases[i] = _lib.GAS[g[0:2]]
        self.plan = plan
        self.nobs = len(obs_mu)
        self.hist_cap = int(hist_cap)
        self.nchains = int(nchains)
        self.ndim = cfg.ndim
        q0a = None if q0 is None else _lib.f64(q0)
        h = ctypes.c_void_p()
        _lib.check(_lib.lib.ngrtd_sampler_create(ctypes.byref(h), ctypes.byref(cfg), plan.handle if plan is not None else None,
                                                 self.nchains, _lib.hptr(q0a), device))
        self.handle = h
        self.device = torch.device("cuda", torch.cuda.current_device() if device < 0 else device)

    def close(self):
        if getattr(self, "handle", None):
            _lib.lib.ngrtd_sampler_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def run(self, nsteps, tune=False, record=False, thin=1, keep_trace=False, stream=None):
        """Advance every chain by nsteps steps (one kernel launch).  Returns the natural-space trace
        [ceil(nsteps/thin), nchains, ndim] as a CUDA tensor if keep_trace, else None."""
        import torch
        trace = None
        if record and keep_trace:
            nd = (int(nsteps) + thin - 1) // thin
            trace = torch.empty((nd, self.nchains, self.ndim), dtype=torch.float64, device=self.device)
        _lib.check(_lib.lib.ngrtd_sampler_run(self.handle, int(nsteps), int(bool(tune)), int(bool(record)), int(thin),
                                              _lib.dptr(trace), _lib.stream_ptr(stream)))
        return trace

    def set_obs_groups(self, obs_mu, obs_sd, chains_per_group):
        """Config 4: obs_mu/obs_sd [ngroups, nobs]; global chains [g*cpg, (g+1)*cpg) are fitted to row g."""
        mu = _lib.f64(np.atleast_2d(obs_mu))
        sd = _lib.f64(np.atleast_2d(obs_sd))
        if mu.shape != sd.shape:
            raise ValueError("obs_mu and obs_sd must have the same shape [ngroups, nobs]")
        if mu.shape[1] != self.nobs:
            raise ValueError("obs_mu / obs_sd need %d columns (one per observation of the sampler), got %d" % (self.nobs, mu.shape[1]))
        _lib.check(_lib.lib.ngrtd_sampler_set_obs_groups(self.handle, _lib.hptr(mu), _lib.hptr(sd), mu.shape[0],
                                                         int(chains_per_group)))

    def set_population(self, chains_per_population):
        """DE-MC-Z with a shared archive per population of consecutive global chains (ter Braak & Vrugt 2008): proposals use
        the histories of all members, as of the start of the current launch -- so run in several launches (e.g. 500 steps
        each) and keep hist_cap >= 2 x steps per launch.  0 = per-chain archives (pymc3's DEMetropolisZ)."""
        _lib.check(_lib.lib.ngrtd_sampler_set_population(self.handle, int(chains_per_population)))

    def stop_tuning(self):
        _lib.check(_lib.lib.ngrtd_sampler_stop_tuning(self.handle))

    _WHAT = {"q": 0, "logp": 1, "lamb": 2, "scaling": 3, "accepted": 4, "mean": 5, "m2": 6, "history": 7, "accepted_window": 8}

    def get(self, what, stream=None):
        import torch
        w = self._WHAT[what]
        shape = (self.nchains, self.ndim) if w in (0, 5, 6) else (self.nchains,)
        if w == 7:
            shape = (self.hist_cap, self.nchains, self.ndim)
        out = torch.empty(shape, dtype=torch.float64, device=self.device)
        _lib.check(_lib.lib.ngrtd_sampler_get(self.handle, w, _lib.dptr(out), _lib.stream_ptr(stream)))
        return out

    def pooled_moments(self, stream=None):
        """K6 on the device: [sum_c mean (nd), sum_c mean^2 (nd), sum_c M2 (nd), chains] over this shard's chains as a CUDA
        tensor of 3*ndim + 1 doubles (ngrtd_sampler_pooled_moments); ranks add these with one all-reduce
        (distributed.pooled_summary)."""
        import torch
        out = torch.empty(3 * self.ndim + 1, dtype=torch.float64, device=self.device)
        _lib.check(_lib.lib.ngrtd_sampler_pooled_moments(self.handle, _lib.dptr(out), _lib.stream_ptr(stream)))
        return out

    def set(self, what, tensor, stream=None):
        _lib.check(_lib.lib.ngrtd_sampler_set(self.handle, self._WHAT[what], _lib.dptr(tensor.contiguous()), _lib.stream_ptr(stream)))

    def info(self):
        a, b, c = ctypes.c_int64(), ctypes.c_int64(), ctypes.c_int64()
        _lib.check(_lib.lib.ngrtd_sampler_info(self.handle, ctypes.byref(a), ctypes.byref(b), ctypes.byref(c)))
        return dict(step=a.value, ndraws=b.value, hist_start=c.value)

    def state_dict(self, history=True):
        """Checkpoint: the complete sampler state as host arrays.  With the history ring the restored
        sampler continues bit-identically; without it (smaller file) DE-MC-Z restarts from an empty history."""
        keys = ["q", "logp", "lamb", "scaling", "accepted", "accepted_window", "mean", "m2"] + (["history"] if history else [])
        d = {k: self.get(k).cpu().numpy() for k in keys}
        d.update(self.info())
        return d

    def load_state_dict(self, d):
        import torch
        for k in ("q", "lamb", "scaling", "accepted", "accepted_window", "mean", "m2", "history", "logp"):   # logp after q
            if k in d:
                t = torch.from_numpy(np.ascontiguousarray(d[k], dtype=np.float64)).to(self.device)
                self.set(k, t)
        hist_start = d["hist_start"] if "history" in d else d["step"]
        _lib.check(_lib.lib.ngrtd_sampler_set_counters(self.handle, int(d["step"]), int(d["ndraws"]), int(hist_start)))
        torch.cuda.synchronize()

    def sample(self, tune, draws, thin=1, keep_trace=True, chunk=None):
        """mc.sample(tune=, draws=, discard_tuned_samples=True): tuning phase, stop_tuning, recorded draws.
        Returns the trace [draws/thin, nchains, ndim] (natural values) or None."""
        import torch
        if tune:
            self.run(tune, tune=True, record=False)
            self.stop_tuning()
        if chunk is None or not keep_trace:
            return self.run(draws, tune=False, record=True, thin=thin, keep_trace=keep_trace)
        parts = []
        done = 0
        while done < draws:
            n = min(chunk - chunk % thin if chunk >= thin else thin, draws - done)
            parts.append(self.run(n, tune=False, record=True, thin=thin, keep_trace=True))
            done += n
        return torch.cat(parts, dim=0)


def philox4x32_10(ctr, key):
    """Device Philox block (known-answer hook)."""
    c = np.ascontiguousarray(ctr, dtype=np.uint32)
    k = np.ascontiguousarray(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    _lib.check(_lib.lib.ngrtd_philox4x32_10(c.ctypes.data_as(ctypes.c_void_p), k.ctypes.data_as(ctypes.c_void_p),
                                            out.ctypes.data_as(ctypes.c_void_p)))
    return out
