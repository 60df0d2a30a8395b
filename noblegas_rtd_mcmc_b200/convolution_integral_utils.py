"""Drop-in for the reference's utils/convolution_integral_utils.py (class tracer_conv_integral), running on the
B200 through libngrtd.so.  Same names, kwargs and mutating-attribute behaviour as the reference
(utils/convolution_integral_utils.py:105-340); additive behaviour: array-valued tau/eta/D/J/lamba of shape [B]
give batched results of shape [B].

'frac_inf_diff' (fracture / matrix diffusion, :36-63,238-266) is evaluated by a dedicated quadrature kernel -- with the
dispersion advective RTD or with a caller-supplied one, `update_pars(f_tadv_ext=...)` (:66-97) -- and convolved through the
external-weights path.  Not provided: 'frac_inf_diff.mint' (an older pure-Python duplicate that prints on every call,
:199-235) and the dead 'SF6' accumulation branch (:330-331, J_sf6 is never set by any caller).

`.g_tp` after convolve() holds what the reference stores there (:300-316): the normalised weights AFTER the decay /
ingrowth factor.  The fused kernel never materialises weights, so the attribute is filled lazily on first access.
"""
import hashlib

import numpy as np

from . import _lib

_MODS = ("piston", "exponential", "exp_pist_flow", "dispersion")


def _as_series(C_t):
    """Return (values newest-first [L], index newest-first [L], last index value) from a DataFrame/Series/array."""
    if hasattr(C_t, "to_numpy") and hasattr(C_t, "index"):
        v = np.asarray(C_t.to_numpy(), dtype=np.float64)
        if v.ndim == 2:
            if v.shape[1] != 1:
                raise ValueError("C_t must hold a single tracer column (reference: 'Input has to be single tracer')")
            v = v[:, 0]
        idx = np.asarray(C_t.index.to_numpy(), dtype=np.float64)
    else:
        v = np.asarray(C_t, dtype=np.float64).ravel()
        idx = np.arange(len(v) - 1, -1, -1, dtype=np.float64)      # same layout as C_in_dict: lag years descending
    return np.ascontiguousarray(v[::-1]), np.ascontiguousarray(idx[::-1]), float(idx[-1])


class tracer_conv_integral():
    def __init__(self, C_t, t_samp):
        self.C_t = C_t            # input series; DataFrame indexed by lag (descending to the sampling date)
        self.t_samp = t_samp      # sample date
        self._plans = {}
        self._g_tp = None         # decayed weights of the last convolve(): ndarray, or a thunk that materialises them

    # ---- .g_tp (reference :300-316: normalised weights times exp(-lam tp), or times 1 - exp(-lam tp) for '3He')
    @property
    def g_tp(self):
        if callable(self._g_tp):
            self._g_tp = self._g_tp()
        return self._g_tp

    @g_tp.setter
    def g_tp(self, value):
        self._g_tp = value

    def _decay(self, g, lam, tp):
        """g * exp(-lam tp), or g * (1 - exp(-lam tp)) for 3He ingrowth (:313-316); lam scalar or [B]."""
        lam = np.asarray(lam, dtype=np.float64)
        e = np.exp(-lam[..., None] * tp) if lam.ndim else np.exp(-lam * tp)
        return g * (1.0 - e) if self.rad_accum == '3He' else g * e

    # ---- utils/convolution_integral_utils.py:111-142
    def update_pars(self, **kwargs):
        self.tau = kwargs.get('tau', None)
        self.mod_type = kwargs.get('mod_type', None)
        self.t_half = kwargs.get('t_half', False)
        if np.any(self.t_half):
            self.thalf_2_lambda(self.t_half)
        else:
            self.lamba = 0.0
        self.rad_accum = kwargs.get('rad_accum', False)
        self.J = kwargs.get('J', False)
        self.eta = kwargs.get('eta', None)
        self.D = kwargs.get('D', None)
        self.bbar = kwargs.get('bbar', None)
        self.Phi_im = kwargs.get('Phi_im', None)
        self.J_sf6 = kwargs.get('J_sf6', None)
        self.f_tadv_ext = kwargs.get('f_tadv_ext', None)

    # ---- :146-152
    def thalf_2_lambda(self, t_half):
        lamba = -1 * np.log(0.5) / t_half
        self.lamba = lamba
        return lamba

    # ------------------------------------------------------------------ helpers
    def _check_model(self):
        if self.mod_type == "frac_inf_diff":
            return
        if self.mod_type not in _MODS:
            if self.mod_type == "frac_inf_diff.mint":
                raise NotImplementedError("mod_type 'frac_inf_diff.mint' is a superseded duplicate; use 'frac_inf_diff'")
            raise ValueError("unknown mod_type %r (known: %s)" % (self.mod_type, ", ".join(_MODS)))

    def _grid(self):
        vals, idx, last = _as_series(self.C_t)
        dtp = float(np.floor(self.t_samp - last))                   # :172
        return vals, idx, dtp

    def _batch(self, *xs):
        n = 1
        scalar = True
        for x in xs:
            if x is not None and x is not False and np.ndim(x) > 0:
                scalar = False
                n = max(n, np.size(x))
        return n, scalar

    @staticmethod
    def _col(x, B, default=0.0):
        if x is None or x is False:
            x = default
        return np.array(np.broadcast_to(np.asarray(x, dtype=np.float64), (B,)), dtype=np.float64)   # writable copy

    # ---- :155-281
    def gen_g_tp(self):
        """Normalised RTD weights: ndarray [L] (or [B, L] for array-valued parameters); sets .tau_list."""
        import torch
        self._check_model()
        vals, _, dtp = self._grid()
        L = len(vals)
        tp = np.arange(0, L).astype(float)
        tp[0] += 1e-5
        tp += dtp
        self.tau_list = tp.copy()
        if self.mod_type == "frac_inf_diff":
            dev = torch.device("cuda")
            fext = getattr(self, "f_tadv_ext", None)
            try:                                                # the reference's test for "an external RTD was given" (:250-252)
                len(fext)
            except TypeError:
                fext = None
            if fext is not None:
                fext = np.ascontiguousarray(fext, dtype=np.float64)
                if fext.ndim != 1 or len(fext) != L:            # the reference only prints a warning (:79-80) and then fails in interp
                    raise ValueError("f_tadv_ext must have the length of the tracer input function (%d)" % L)
                B, scalar = self._batch(self.bbar, self.Phi_im)
                t = [torch.from_numpy(self._col(v, B)).to(dev) for v in (self.bbar, self.Phi_im)]
                fe = torch.from_numpy(fext).to(dev)
            else:
                B, scalar = self._batch(self.tau, self.D, self.bbar, self.Phi_im)
                t = [torch.from_numpy(self._col(v, B)).to(dev) for v in (self.tau, self.D, self.bbar, self.Phi_im)]
            g = torch.empty((B, L), dtype=torch.float64, device=dev)
            mu = torch.empty(B, dtype=torch.float64, device=dev)
            if fext is not None:
                _lib.check(_lib.lib.ngrtd_rtd_weights_fdm_ext_dev(L, dtp, _lib.dptr(fe), _lib.dptr(t[0]), _lib.dptr(t[1]), B,
                                                                  _lib.dptr(g), _lib.dptr(mu), _lib.stream_ptr()))
            else:
                _lib.check(_lib.lib.ngrtd_rtd_weights_fdm_dev(L, dtp, _lib.dptr(t[0]), _lib.dptr(t[1]), _lib.dptr(t[2]), _lib.dptr(t[3]),
                                                              B, _lib.dptr(g), _lib.dptr(mu), _lib.stream_ptr()))
            out, m = g.cpu().numpy(), mu.cpu().numpy()
            self.FM_mu = float(m[0]) if scalar else m          # mean travel time (:264-265)
            return out[0] if scalar else out
        B, scalar = self._batch(self.tau, self.eta, self.D)
        dev = torch.device("cuda")
        tau = torch.from_numpy(self._col(self.tau, B)).to(dev)
        eta = torch.from_numpy(self._col(self.eta, B, 1.0)).to(dev) if self.mod_type == "exp_pist_flow" else None
        D = torch.from_numpy(self._col(self.D, B)).to(dev) if self.mod_type == "dispersion" else None
        g = torch.empty((B, L), dtype=torch.float64, device=dev)
        _lib.check(_lib.lib.ngrtd_rtd_weights_dev(_lib.MOD[self.mod_type], L, dtp, _lib.dptr(tau), _lib.dptr(eta),
                                                  _lib.dptr(D), B, _lib.dptr(g), _lib.stream_ptr()))
        out = g.cpu().numpy()
        return out[0] if scalar else out

    # ---- :284-340
    def convolve(self, **kwargs):
        """Concentration at the sampling date: float (or ndarray [B])."""
        self._check_model()
        if self.rad_accum == 'SF6':
            raise ValueError("rad_accum='SF6' needs J_sf6, which no caller of the reference sets (dead branch)")
        if self.rad_accum not in (False, None, '3He', '4He'):
            raise ValueError("unknown rad_accum %r" % (self.rad_accum,))
        g_tau = kwargs.get('g_tau', None)
        vals, idx, dtp = self._grid()
        L = len(vals)
        if g_tau is None and self.mod_type == "frac_inf_diff":
            g_tau = self.gen_g_tp()                             # weights by quadrature, then the external-weights path
        if g_tau is not None:
            return self._convolve_external(np.asarray(g_tau, dtype=np.float64), vals, idx, dtp)
        lam = getattr(self, "lamba", 0.0)
        B, scalar = self._batch(self.tau, self.eta, self.D, self.J, lam)
        lam_batched = np.ndim(lam) > 0
        ra = self.rad_accum if self.rad_accum else False
        if lam_batched and ra:
            raise ValueError("array-valued lamba together with rad_accum is not supported")
        # the reference re-reads C_t on every convolve(): key the plan cache on the CONTENT of the series and its index
        digest = hashlib.blake2b(vals.tobytes() + idx.tobytes(), digest_size=16).digest()
        key = (self.mod_type, ra, None if lam_batched else float(lam), digest, L, dtp)
        plan = self._plans.get(key)
        if plan is None:
            desc = dict(series=0, rad_accum=ra, lam=0.0 if lam_batched else float(lam), use_thalf_cfc=lam_batched)
            plan = _lib.Plan(vals.reshape(-1, 1), [desc], self.mod_type, False, lag_index=idx, dtp=dtp)
            if len(self._plans) >= 8:          # callers toggle mod_type/tau between convolve() calls (run_age_mcmc.py:299-303)
                self._plans.pop(next(iter(self._plans)))
            self._plans[key] = plan
        names, cols = ["tau1"], [self._col(self.tau, B)]
        if self.mod_type == "exp_pist_flow":
            names.append("eta1"); cols.append(self._col(self.eta, B))
        if self.mod_type == "dispersion":
            names.append("D1"); cols.append(self._col(self.D, B))
        if ra == '4He':
            with np.errstate(divide="ignore"):
                names.append("J"); cols.append(np.log10(self._col(self.J, B)))
        if lam_batched:
            with np.errstate(divide="ignore"):
                names.append("thalf_cfc"); cols.append(np.log(2.0) / self._col(lam, B))
        theta = np.ascontiguousarray(np.stack(cols, axis=1))
        out = plan.forward_host(theta, names)[:, 0]
        # weights are generated in registers by the fused kernel; .g_tp materialises them (decay applied) on first access
        state = (self.mod_type, self.tau, self.eta, self.D, lam, self.rad_accum)

        def materialise():
            keep = (self.mod_type, self.tau, self.eta, self.D)
            self.mod_type, self.tau, self.eta, self.D = state[:4]
            try:
                g = self.gen_g_tp()
            finally:
                self.mod_type, self.tau, self.eta, self.D = keep
            keep_ra, self.rad_accum = self.rad_accum, state[5]
            try:
                return self._decay(g, state[4], self.tau_list)
            finally:
                self.rad_accum = keep_ra
        self._g_tp = materialise
        C_i = float(out[0]) if scalar else out
        self.C_i = C_i
        return C_i

    def _convolve_external(self, g, vals, idx, dtp):
        import torch
        scalar = g.ndim == 1
        g2 = np.ascontiguousarray(np.atleast_2d(g))
        B, L = g2.shape
        if L != len(vals):
            raise ValueError("g_tau has %d lags, the input series %d" % (L, len(vals)))
        dev = torch.device("cuda")
        lam = getattr(self, "lamba", 0.0)
        ra = _lib.ACC[self.rad_accum if self.rad_accum else False]
        gt = torch.from_numpy(g2).to(dev)
        st = torch.from_numpy(vals).to(dev)
        it = torch.from_numpy(idx).to(dev)
        lt = torch.from_numpy(self._col(lam, B)).to(dev)
        jt = torch.from_numpy(self._col(self.J, B)).to(dev) if ra == 2 else None
        out = torch.empty(B, dtype=torch.float64, device=dev)
        _lib.check(_lib.lib.ngrtd_convolve_g_dev(L, dtp, _lib.dptr(gt), B, _lib.dptr(st), _lib.dptr(it), _lib.dptr(lt), ra,
                                                 _lib.dptr(jt), _lib.dptr(out), _lib.stream_ptr()))
        res = out.cpu().numpy()
        tp = np.arange(0, L).astype(float)
        tp[0] += 1e-5
        tp += dtp
        self._g_tp = self._decay(g, lam if np.ndim(lam) else float(lam), tp)      # decayed, as the reference stores it
        self.C_i = float(res[0]) if scalar else res
        return self.C_i
