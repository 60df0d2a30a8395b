"""Operator boundary of the age model: drop-in for `ForwardMod` of the reference's
age_ens_runs_mcmc/run_age_mcmc_utils.py:47-163, plus the batched joint operator the device sampler uses.

`ForwardMod(conv_kwgs, par_names, tracer)` keeps the reference constructor and `perform(node, inputs, outputs)`
(one theta -> one scalar); `perform_batch(theta[B, ndim]) -> out[B]` is the additive batched form.
`JointForwardMod(conv_kwgs_by_tracer, par_names, tracers)` evaluates every tracer of a joint inversion in ONE
kernel launch (the RTD of a chain is generated once and shared by its tracers).
"""
import numpy as np

from . import _lib
from .convolution_integral_utils import _as_series

P_NAMES = ("tau1", "tau2", "f1", "f2", "eta1", "eta2", "D1", "D2", "J", "thalf_cfc", "lamsf6")


def _tracer_desc(kw, tracer, par_names, series_col):
    t_half = kw.get('t_half', False)
    ra = kw.get('rad_accum', False)
    if ra not in (False, None, '3He', '4He'):
        raise ValueError("unknown rad_accum %r for tracer %s" % (ra, tracer))
    use_cfc = ('thalf_cfc' in par_names and tracer == 'CFC12')          # run_age_mcmc_utils.py:107-109
    if use_cfc and ra:
        raise ValueError("thalf_cfc on a rad_accum tracer is not supported")
    return dict(series=series_col, rad_accum=ra if ra else False,
                lam=float(-1 * np.log(0.5) / t_half) if t_half else 0.0,   # update_pars -> thalf_2_lambda
                use_thalf_cfc=use_cfc, use_lamsf6=(tracer == 'SF6'))       # :160-161


class JointForwardMod(object):
    """All tracers of one well/model pair in one plan (input series resident in HBM)."""

    def __init__(self, conv_kwgs, par_names, tracers, device=-1):
        self.p_names = list(par_names)
        self.tracers = list(tracers)
        for p in self.p_names:
            if p not in P_NAMES:
                raise ValueError("unknown parameter %r (known: %s)" % (p, ", ".join(P_NAMES)))
        mod1 = mod2 = None
        cols, descs, lag_index, dtp, L = [], [], None, None, None
        for t in self.tracers:
            kw = conv_kwgs[t]
            if kw.get('bbar', False) or kw.get('Phi_im', False):
                raise NotImplementedError("fracture/matrix-diffusion parameters are outside the B200 hot path")
            m1 = kw.get('mod_type1', conv_kwgs.get('mod_type1', False))
            m2 = kw.get('mod_type2', conv_kwgs.get('mod_type2', False))
            if mod1 is None:
                mod1, mod2 = m1, m2
            elif (m1, m2) != (mod1, mod2):
                raise ValueError("all tracers of a joint operator share mod_type1/mod_type2")
            vals, idx, last = _as_series(kw['C_t'])
            d = float(np.floor(last - last))                              # t_samp = C_t.index[-1] (:94)
            if L is None:
                L, lag_index, dtp = len(vals), idx, d
            elif len(vals) != L or not np.array_equal(idx, lag_index):
                raise ValueError("all tracers of a joint operator share the lag grid")
            col = -1
            if np.any(vals != 0.0):
                for i, c in enumerate(cols):
                    if np.array_equal(c, vals):
                        col = i
                        break
                else:
                    cols.append(vals)
                    col = len(cols) - 1
            descs.append(_tracer_desc(kw, t, self.p_names, col))
        X = np.stack(cols, axis=1) if cols else np.zeros((L, 1))
        self.plan = _lib.Plan(X, descs, mod1, mod2 if mod2 else False, lag_index=lag_index, dtp=dtp, device=device)
        self.mod_type1, self.mod_type2 = mod1, mod2

    def perform_batch(self, theta):
        """theta [B, ndim] ordered by par_names -> modelled concentrations [B, ntracer] (host numpy)."""
        return self.plan.forward_host(np.atleast_2d(theta), self.p_names)

    def perform_batch_device(self, theta_t, out_t=None, stream=None):
        """torch CUDA tensor in / out, no synchronisation."""
        return self.plan.forward_dev(theta_t, self.p_names, out_t, stream)

    def logp_batch(self, theta, obs_mu, obs_err, kind="studentt", nu=None):
        return self.plan.forward_loglik_host(np.atleast_2d(theta), self.p_names, obs_mu, obs_err, kind, nu)


class ForwardMod(object):
    '''Single-tracer operator with the reference's constructor and perform() (Theano Op protocol).'''

    def __init__(self, conv_kwgs, par_names, tracer):
        kwargs = conv_kwgs.copy()
        self.C_t = kwargs.get('C_t', False)
        self.mod_type1 = kwargs.get('mod_type1', False)
        self.mod_type2 = kwargs.get('mod_type2', False)
        self.t_half = kwargs.get('t_half', False)
        self.rad_accum = kwargs.get('rad_accum', False)
        self.J = kwargs.get('J', False)
        self.eta = kwargs.get('eta', False)
        self.D = kwargs.get('D', False)
        self.bbar = kwargs.get('bbar', False)
        self.Phi_im = kwargs.get('Phi_im', False)
        self.t = tracer
        self.p_names = par_names
        self._joint = JointForwardMod({tracer: kwargs}, par_names, [tracer])

    def perform_batch(self, theta):
        return self._joint.perform_batch(theta)[:, 0]

    def perform(self, node, inputs, outputs):
        """Forward model for one parameter vector (inputs[0]), result in outputs[0][0] as in the reference."""
        outputs[0][0] = np.array(self.perform_batch(np.asarray(inputs[0], dtype=np.float64).reshape(1, -1))[0])

    def as_theano_op(self):
        """The reference's boundary is `class ForwardMod(TT.Op)` with `itypes = [TT.dvector]`, `otypes = [TT.dscalar]`
        (run_age_mcmc_utils.py:47-52), instantiated once per tracer inside the pymc3 model (:376-383).  Under a real
        pymc3 / Theano installation this returns exactly such an Op whose `perform` is this object's `perform`, so
        `ForwardMod(ckw, par_names, t).as_theano_op()(theta)` drops into `build_mcmc_model_joint_sT` unchanged.
        (pymc3 evaluates one theta per call: the batched device sampler, `conv_mcmc.sample_mcmc`, is the fast path.)"""
        return as_theano_op(self)


def as_theano_op(forward_mod):
    """Wrap a ForwardMod-like object (anything with `perform(node, inputs, outputs)`) as a Theano / Aesara Op with the
    reference's signature: dvector -> dscalar.  Raises ImportError when neither theano nor aesara is installed (this image)."""
    try:
        import theano.tensor as TT
    except ImportError:
        import aesara.tensor as TT          # pymc3 >= 3.11.5 / pymc 4 successor of theano

    class _ForwardModOp(TT.Op):
        itypes = [TT.dvector]
        otypes = [TT.dscalar]

        def perform(self, node, inputs, outputs):
            forward_mod.perform(node, inputs, outputs)
    return _ForwardModOp()


# ---------------------------------------------------------------------------------------------------------------
# Model builder + sampler driver: the role of conv_mcmc (run_age_mcmc_utils.py:189-429 of the reference)
# ---------------------------------------------------------------------------------------------------------------
class conv_mcmc(object):
    """Same constructor as the reference; priors, observation errors, likelihood and sampler settings follow
    build_mcmc_model_joint_sT / sample_mcmc (:275-429).  With `savedir` the finished trace is written where and how the
    reference writes it (`./<savedir>/<well>.<tracers>.<model>.<savenum>.netcdf`, :242-256, az.to_netcdf :425) as a
    NetCDF-4 file the reference's plotting scripts open with az.from_netcdf (:434, post_plots.py:119-148).  Plotting is
    out of scope."""

    def __init__(self, well, tracer, obs_kwgs, conv_kwgs, prior_kwgs, savedir=None, savenum=None):
        self.well = well
        self.tracer = list(tracer)
        g = prior_kwgs.get
        self.par_names = list(g('par_names', []))
        self.obs = obs_kwgs
        self.conv_kwgs = conv_kwgs
        self.mod_type1 = conv_kwgs.get('mod_type1', False)
        self.mod_type2 = conv_kwgs.get('mod_type2', False)
        self.prior_kwgs = dict(prior_kwgs)
        self.savenum = savenum
        self.savedir = savedir
        self.idata = None
        self.trace_name = None
        if savedir is not None:                                                  # setup_dirs (:242-256)
            import os
            tracer_list = '.'.join(self.tracer)
            mod_type = '{}-{}'.format(self.mod_type1, self.mod_type2) if self.mod_type2 else '{}'.format(self.mod_type1)
            self.trace_name = os.path.join('.', str(savedir), '{}.{}.{}.{}.netcdf'.format(self.well, tracer_list, mod_type, savenum))

    # ---- :286-344
    def build_priors(self):
        from .sampler import prior
        pk = self.prior_kwgs
        pri = [prior("uniform", "tau1", pk['tau1_low'], pk['tau1_high']),
               prior("beta", "nu_", 2.0, 0.1, 0.0, 1.0)]                       # nu = nu_*(30-5)+5   (:291-292)
        if 'J' in self.par_names:
            pri.append(prior("normal", "J", pk['J_mu'], pk['J_sd']))
        if self.mod_type2:
            pri.append(prior("uniform", "tau2", pk['tau2_low'], pk['tau2_high']))
            pri.append(prior("uniform", "f1", pk['f1_low'], pk['f1_high']))
        if self.mod_type1 == 'exp_pist_flow':
            pri.append(prior("uniform", "eta1", pk['eta1_low'], pk['eta1_high']))
        if self.mod_type2 == 'exp_pist_flow':
            pri.append(prior("uniform", "eta2", pk['eta2_low'], pk['eta2_high']))
        if self.mod_type1 == 'dispersion':
            pri.append(prior("uniform", "D1", pk['D1_low'], pk['D1_high']))
        if self.mod_type2 == 'dispersion':
            pri.append(prior("uniform", "D2", pk['D2_low'], pk['D2_high']))
        if 'thalf_cfc' in self.par_names:
            pri.append(prior("beta", "thalf_cfc", 2.0, 2.0, pk['cfc_thalf_lo'], pk['cfc_thalf_hi']))
        if 'lamsf6' in self.par_names:
            pri.append(prior("halfnormal", "lamsf6", 0.5 / 3))
        return pri

    # ---- :353-356
    def observations(self):
        mu = np.array([np.asarray(self.obs[w]['obs_df'], dtype=np.float64).mean() for w in self.tracer])
        nerr = np.array([np.asarray(self.obs[w]['obs_df'], dtype=np.float64).std() for w in self.tracer])
        perr = mu * np.array([self.obs[w]['obs_perr'] for w in self.tracer])
        return mu, nerr + perr

    def sample_mcmc(self, chains=3, tune=10000, draws=10000, random_seed=123423, tune_interval=1000, thin=1,
                    likelihood="studentt", hist_cap=None):
        """mc.DEMetropolisZ(tune_interval=1000); mc.sample(tune=10000, draws=10000, chains=3, ...) (:412-417).
        Returns {'posterior': {var: ndarray[chain, draw]}, 'sample_stats': {...}}."""
        ckw = {t: dict(self.conv_kwgs[t], mod_type1=self.mod_type1, mod_type2=self.mod_type2) for t in self.tracer}
        joint = JointForwardMod(ckw, self.par_names, self.tracer)
        from .sampler import Sampler
        pri = self.build_priors()
        mu, err = self.observations()
        smp = Sampler(pri, mu, err, chains, plan=joint.plan, lik=likelihood, nu_range=(5.0, 30.0),
                      f2_from_f1=bool(self.mod_type2), tune_interval=tune_interval,
                      hist_cap=hist_cap or (tune + draws), seed=random_seed)
        import time as _time
        import torch as _torch
        _torch.cuda.synchronize()
        t0 = _time.perf_counter()
        trace_t = smp.sample(tune, draws, thin=thin)
        _torch.cuda.synchronize()
        self.sampling_time = _time.perf_counter() - t0                       # what pymc3 stores as `sampling_time`
        trace = trace_t.cpu().numpy()                                        # [draw, chain, dim]
        post = {n: trace[:, :, i].T.copy() for i, n in enumerate(smp.names)}
        post['nu'] = post['nu_'] * (30.0 - 5.0) + 5.0
        if self.mod_type2:
            post['f2'] = 1.0 - post['f1']
            post['tau'] = post['f1'] * post['tau1'] + post['f2'] * post['tau2']          # :305
        self.idata = {'posterior': post,
                      'sample_stats': {'accept_rate': smp.get("accepted").cpu().numpy() / float(tune + draws),
                                       'lamb': smp.get("lamb").cpu().numpy(), 'sampling_time': self.sampling_time},
                      'observed_data': {"obs_%s" % t: np.array([m]) for t, m in zip(self.tracer, mu)}}
        if self.trace_name is not None:                                          # az.to_netcdf(idata, self.trace_name), :425
            import os
            from . import diagnostics
            os.makedirs(os.path.dirname(self.trace_name), exist_ok=True)
            diagnostics.save_trace(self.trace_name, post, observed_data=self.idata['observed_data'],
                                   attrs={"sampling_time": np.array([self.sampling_time]), "tuning_steps": np.array([tune]),
                                          "inference_library": "ngrtd-b200 (DEMetropolisZ, pymc3 3.11.2 semantics)"})
        return self.idata

    def posterior_predictive(self, idata=None, tracers=None, chain=None, max_draws=None):
        """Batched replacement of the reference's posterior-predictive Python loop (run_age_mcmc.py:243-319): one
        launch evaluates every posterior draw for every tracer.  Returns {tracer: ndarray[draws]}."""
        idata = idata or self.idata
        post = idata['posterior']
        tracers = list(tracers or self.tracer)
        ckw = {t: dict(self.conv_kwgs[t], mod_type1=self.mod_type1, mod_type2=self.mod_type2) for t in tracers}
        joint = JointForwardMod(ckw, self.par_names, tracers)
        sel = slice(None) if chain is None else chain                      # the reference uses chain 0 only (:277)
        cols = [np.asarray(post[p])[sel].reshape(-1) for p in self.par_names]
        theta = np.ascontiguousarray(np.stack(cols, axis=1))
        if max_draws is not None:
            theta = theta[:max_draws]
        out = joint.perform_batch(theta)
        return {t: out[:, i].copy() for i, t in enumerate(tracers)}
