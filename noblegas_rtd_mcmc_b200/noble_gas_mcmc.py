"""Operator boundary of the noble-gas fit: batched `ce_exc_wrapper` of the reference's
ng_interp/noble_gas_mcmc.py:205-213 (theta = [log10 Ae, log10 F, E, T] -> modelled Ne/Ar/Kr/Xe)."""
import numpy as np

from . import _lib

GASES = ['Ne', 'Ar', 'Kr', 'Xe']     # ng_interp/noble_gas_mcmc.py:107


def ce_exc_wrapper(theta, gases=GASES):
    """theta [4] -> ndarray [len(gases)] (reference signature), or theta [B, 4] -> [B, len(gases)]."""
    import torch
    th = np.asarray(theta, dtype=np.float64)
    single = th.ndim == 1
    th = np.ascontiguousarray(np.atleast_2d(th))
    if th.shape[1] != 4:
        raise ValueError("theta must be [log10 Ae, log10 F, E, T]")
    ids = _lib.i32([_lib.GAS[g[0:2]] for g in gases])
    t = torch.from_numpy(th).cuda()
    out = torch.empty((th.shape[0], len(gases)), dtype=torch.float64, device=t.device)
    _lib.check(_lib.lib.ngrtd_ce_wrapper_dev(len(gases), _lib.hptr(ids), _lib.dptr(t), th.shape[0], _lib.dptr(out),
                                             _lib.stream_ptr()))
    res = out.cpu().numpy()
    return res[0] if single else res


def as_theano_op(gases=GASES):
    """`ce_exc_wrapper` as the Theano Op the reference builds with `@as_op(itypes=[tt.dvector], otypes=[tt.dvector])`
    (ng_interp/noble_gas_mcmc.py:205): dvector theta -> dvector of modelled gases.  Needs theano (or aesara); raises
    ImportError otherwise (this image)."""
    try:
        import theano.tensor as tt
        from theano.compile.ops import as_op
    except ImportError:
        import aesara.tensor as tt
        from aesara.compile.ops import as_op

    @as_op(itypes=[tt.dvector], otypes=[tt.dvector])
    def _op(theta):
        return np.asarray(ce_exc_wrapper(theta, gases), dtype=np.float64)
    return _op


class mcmc_model(object):
    """The noble-gas closed-equilibrium inversion of ng_interp/noble_gas_mcmc.py:95-267 (priors :118-147,224-250;
    Student-T likelihood :254-266) with the reference's sampler settings (:408-415) as defaults."""

    well_elev = {'PLM1': 2786.889893, 'PLM6': 2759.569824, 'PLM7': 2782.550049}     # :80-82

    def __init__(self, obs_dict, well_elev, err_dict=None, gases=GASES, lapse_slope=-146.0, err_lapse_slope=17.0,
                 lapse_b=3354.0, Emax=3300.0):
        self.gases = list(gases)
        self.err_dict = err_dict or {'He': 1.5, 'Ne': 1.5, 'Ar': 2.5, 'Kr': 3.1, 'Xe': 15.1}     # :104
        self.obs_mu = np.array([obs_dict[g] for g in self.gases], dtype=np.float64)
        self.obs_sd = self.obs_mu * np.array([self.err_dict[g] / 100 for g in self.gases])        # :254-256
        err_b = abs(lapse_slope) * 2.5
        self.par_bnd = {'Ae': (-4.0, -1.0), 'F': (-1.0, 1.0), 'E': (well_elev - 10.0, Emax),
                        'b': (lapse_b - err_b, lapse_b + err_b)}                                  # :118-147
        self.lapse_slope, self.err_lapse_slope = lapse_slope, err_lapse_slope

    def build_priors(self):
        from .sampler import prior
        b = self.par_bnd
        return [prior("beta", "log10Ae", 2, 2, *b['Ae']), prior("beta", "log10F", 2, 2, *b['F']),
                prior("beta", "E", 2, 4, *b['E']), prior("normal", "m", self.lapse_slope, self.err_lapse_slope),
                prior("beta", "b", 2, 2.5, *b['b']), prior("beta", "nu_", 2.0, 0.1, 0.0, 1.0)]

    def sample(self, chains=4, tune=10000, draws=50000, random_seed=123423, tune_interval=5000, thin=1, hist_cap=None,
               trace_path=None):
        """trace_path: write the transformed posterior as NetCDF-4, the file the reference saves as
        `traces/<well>_trans.netcdf` (noble_gas_mcmc.py:438-447) and prep.py / compplots.py open with az.from_netcdf."""
        from .sampler import Sampler
        smp = Sampler(self.build_priors(), self.obs_mu, self.obs_sd, chains, plan=None, gases=self.gases,
                      lik="studentt", nu_range=(1.0, 30.0), tune_interval=tune_interval,
                      hist_cap=hist_cap or (tune + draws), seed=random_seed)
        trace = smp.sample(tune, draws, thin=thin).cpu().numpy()
        raw = {n: trace[:, :, i].T.copy() for i, n in enumerate(smp.names)}
        b = self.par_bnd
        post = {'m': raw['m'], 'b': raw['b'], 'E': raw['E'], 'nu_': raw['nu_'], 'nu': raw['nu_'] * 29.0 + 1.0,
                'Ae': 10 ** raw['log10Ae'], 'F': 10 ** raw['log10F'],                             # :438-439
                'Ae_beta': (raw['log10Ae'] - b['Ae'][0]) / (b['Ae'][1] - b['Ae'][0]),
                'F_beta': (raw['log10F'] - b['F'][0]) / (b['F'][1] - b['F'][0]),
                'E_beta': (raw['E'] - b['E'][0]) / (b['E'][1] - b['E'][0]),
                'b_beta': (raw['b'] - b['b'][0]) / (b['b'][1] - b['b'][0])}
        post['T'] = (post['E'] - post['b']) / post['m']                                           # :240
        self.sampler = smp
        if trace_path is not None:
            from . import diagnostics
            diagnostics.save_trace(trace_path, post, observed_data={g: np.array([v]) for g, v in zip(self.gases, self.obs_mu)},
                                   attrs={"tuning_steps": np.array([tune]), "inference_library": "ngrtd-b200 (DEMetropolisZ, pymc3 3.11.2 semantics)"})
        return {'posterior': post, 'sample_stats': {'accept_rate': smp.get("accepted").cpu().numpy() / float(tune + draws)}}
