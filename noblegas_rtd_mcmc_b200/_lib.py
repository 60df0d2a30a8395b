"""ctypes binding of libngrtd.so (C ABI declared in include/ngrtd.h).

There is NO CPU fallback: importing this module without the compiled CUDA library raises.
PyTorch is used by the callers only for device memory and streams (tensor.data_ptr(),
torch.cuda.current_stream().cuda_stream); no torch type crosses this boundary.
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("NGRTD_LIB", os.path.join(_HERE, "libngrtd.so"))

MOD = {"piston": 1, "exponential": 2, "exp_pist_flow": 3, "dispersion": 4}
# "f1_f2c": one column that is f1 and makes f2 = 1 - f1 on the device (NGRTD_P_F1_COMPLEMENT; the reference's model has
# f2 = Deterministic(1 - f1), run_age_mcmc_utils.py:304) -- for batched callers that want to drop the redundant column
SLOT = {"tau1": 0, "tau2": 1, "f1": 2, "f2": 3, "eta1": 4, "eta2": 5, "D1": 6, "D2": 7, "J": 8,
        "thalf_cfc": 9, "lamsf6": 10, "f1_f2c": 11}
ACC = {False: 0, None: 0, "3He": 1, "4He": 2}
GAS = {"He": 0, "Ne": 1, "Ar": 2, "Kr": 3, "Xe": 4}
LIK = {"normal": 0, "studentt": 1}


class NgrtdError(RuntimeError):
    pass


class Tracer(ctypes.Structure):
    _fields_ = [("series", ctypes.c_int32), ("rad_accum", ctypes.c_int32), ("lam", ctypes.c_double),
                ("use_thalf_cfc", ctypes.c_int32), ("use_lamsf6", ctypes.c_int32)]


class Prior(ctypes.Structure):
    _fields_ = [("kind", ctypes.c_int32), ("target", ctypes.c_int32), ("p0", ctypes.c_double), ("p1", ctypes.c_double),
                ("lo", ctypes.c_double), ("hi", ctypes.c_double)]


class SamplerCfg(ctypes.Structure):
    _fields_ = [("ndim", ctypes.c_int32), ("prior", Prior * 10), ("lik_kind", ctypes.c_int32), ("nu_sampled", ctypes.c_int32),
                ("nu_lo", ctypes.c_double), ("nu_hi", ctypes.c_double), ("nu_fixed", ctypes.c_double),
                ("nobs", ctypes.c_int32), ("obs_mu", ctypes.c_double * 8), ("obs_sd", ctypes.c_double * 8),
                ("f2_from_f1", ctypes.c_int32), ("proposal_dist", ctypes.c_int32), ("de_mcz", ctypes.c_int32),
                ("tune_target", ctypes.c_int32), ("tune_interval", ctypes.c_int32), ("scaling", ctypes.c_double),
                ("lamb", ctypes.c_double), ("tune_drop_fraction", ctypes.c_double), ("hist_cap", ctypes.c_int32),
                ("seed", ctypes.c_uint64), ("chain_offset", ctypes.c_int64), ("ngas", ctypes.c_int32),
                ("gases", ctypes.c_int32 * 5)]


PRIOR_KIND = {"uniform": 0, "beta": 1, "normal": 2, "halfnormal": 3}
VAL_NU = 11
NG_TARGET = {"log10Ae": 0, "log10F": 1, "E": 2, "m": 3, "b": 4, "nu_": VAL_NU}

if not os.path.exists(LIB_PATH):
    raise ImportError(
        "libngrtd.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
        "or noblegas_rtd_mcmc_b200/build.py; there is no CPU fallback." % LIB_PATH)

lib = ctypes.CDLL(LIB_PATH)

_vp, _i32, _i64, _dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
_PROTOS = {
    "ngrtd_version": ([], ctypes.c_int),
    "ngrtd_build_features": ([], ctypes.c_int),
    "ngrtd_host_alloc": ([ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t, ctypes.c_int32], ctypes.c_int),
    "ngrtd_host_free": ([ctypes.c_void_p], ctypes.c_int),
    "ngrtd_fp64_peak_probe": ([ctypes.c_int32, ctypes.c_int32, ctypes.POINTER(ctypes.c_double)], ctypes.c_int),
    "ngrtd_last_error": ([], ctypes.c_char_p),
    "ngrtd_plan_create": ([ctypes.POINTER(_vp), _i32, _i32, _vp, _vp, _dbl, _i32, ctypes.POINTER(Tracer), _i32, _i32, _i32], ctypes.c_int),
    "ngrtd_plan_destroy": ([_vp], ctypes.c_int),
    "ngrtd_plan_ntracer": ([_vp], ctypes.c_int),
    "ngrtd_forward_dev": ([_vp, _vp, _i64, _i32, _vp, _vp, _vp], ctypes.c_int),
    "ngrtd_forward_host": ([_vp, _vp, _i64, _i32, _vp, _vp], ctypes.c_int),
    "ngrtd_forward_loglik_dev": ([_vp, _vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _vp], ctypes.c_int),
    "ngrtd_forward_loglik_host": ([_vp, _vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp], ctypes.c_int),
    "ngrtd_forward_loglik_host_submit": ([_vp, _vp, _i64, _i32, _vp, _i32, _vp, _vp, _vp, _vp, _vp, _i32], ctypes.c_int),
    "ngrtd_host_wait": ([_vp, _i32], ctypes.c_int),
    "ngrtd_rtd_weights_dev": ([_i32, _i32, _dbl, _vp, _vp, _vp, _i64, _vp, _vp], ctypes.c_int),
    "ngrtd_rtd_weights_fdm_dev": ([_i32, _dbl, _vp, _vp, _vp, _vp, _i64, _vp, _vp, _vp], ctypes.c_int),
    "ngrtd_rtd_weights_fdm_ext_dev": ([_i32, _dbl, _vp, _vp, _vp, _i64, _vp, _vp, _vp], ctypes.c_int),
    "ngrtd_convolve_g_dev": ([_i32, _dbl, _vp, _i64, _vp, _vp, _vp, _i32, _vp, _vp, _vp], ctypes.c_int),
    "ngrtd_ce_dev": ([_i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _dbl, _i64, _vp, _vp], ctypes.c_int),
    "ngrtd_ce_host": ([_i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _dbl, _i64, _vp], ctypes.c_int),
    "ngrtd_ce_wrapper_dev": ([_i32, _vp, _vp, _i64, _vp, _vp], ctypes.c_int),
    "ngrtd_loglik_dev": ([_i32, _i32, _vp, _vp, _vp, _vp, _i64, _vp, _vp], ctypes.c_int),
    "ngrtd_cfc_dev": ([_i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _dbl, _i64, _vp, _vp], ctypes.c_int),
    "ngrtd_cfc_host": ([_i32, _i32, _vp, _vp, _vp, _vp, _vp, _vp, _dbl, _i64, _vp], ctypes.c_int),
    "ngrtd_sampler_create": ([ctypes.POINTER(_vp), ctypes.POINTER(SamplerCfg), _vp, _i64, _vp, _i32], ctypes.c_int),
    "ngrtd_sampler_destroy": ([_vp], ctypes.c_int),
    "ngrtd_sampler_set_obs_groups": ([_vp, _vp, _vp, _i64, _i64], ctypes.c_int),
    "ngrtd_sampler_set_population": ([_vp, _i64], ctypes.c_int),
    "ngrtd_sampler_run": ([_vp, _i64, _i32, _i32, _i32, _vp, _vp], ctypes.c_int),
    "ngrtd_sampler_stop_tuning": ([_vp], ctypes.c_int),
    "ngrtd_sampler_get": ([_vp, _i32, _vp, _vp], ctypes.c_int),
    "ngrtd_sampler_set": ([_vp, _i32, _vp, _vp], ctypes.c_int),
    "ngrtd_sampler_set_counters": ([_vp, _i64, _i64, _i64], ctypes.c_int),
    "ngrtd_sampler_pooled_moments": ([_vp, _vp, _vp], ctypes.c_int),
    "ngrtd_sampler_info": ([_vp, ctypes.POINTER(_i64), ctypes.POINTER(_i64), ctypes.POINTER(_i64)], ctypes.c_int),
    "ngrtd_philox4x32_10": ([_vp, _vp, _vp], ctypes.c_int),
}
for _name, (_args, _res) in _PROTOS.items():
    _fn = getattr(lib, _name)
    _fn.argtypes = _args
    _fn.restype = _res

EXPORTED = tuple(_PROTOS)


def check(rc):
    if rc != 0:
        raise NgrtdError("libngrtd error %d: %s" % (rc, lib.ngrtd_last_error().decode()))


def fp64_peak_probe(device=-1):
    """measured FP64 peak of a device: {"dfma": TFLOP/s, "dmma": TFLOP/s} (ngrtd_fp64_peak_probe)"""
    out = {}
    for name, kind in (("dfma", 0), ("dmma", 1)):
        v = ctypes.c_double()
        check(lib.ngrtd_fp64_peak_probe(int(device), kind, ctypes.byref(v)))
        out[name] = v.value
    return out


class _HostBlock(object):
    """owner of one ngrtd_host_alloc block (freed with the last numpy view)"""

    def __init__(self, nbytes, write_combined):
        self.p = ctypes.c_void_p()
        check(lib.ngrtd_host_alloc(ctypes.byref(self.p), nbytes, 1 if write_combined else 0))

    def __del__(self):
        try:
            if self.p:
                lib.ngrtd_host_free(self.p)
        except Exception:
            pass


def host_array(shape, write_combined=False):
    """float64 ndarray in page-locked host memory (ngrtd_host_alloc) for the *_host calls; write_combined=True for buffers
    the CPU only writes (theta batches): faster on the way to the device, slow to read back on the CPU."""
    shape = tuple(int(x) for x in (shape if isinstance(shape, (tuple, list)) else (shape,)))
    n = int(np.prod(shape))
    blk = _HostBlock(max(n, 1) * 8, write_combined)
    buf = (ctypes.c_double * max(n, 1)).from_address(blk.p.value)
    buf._ngrtd_owner = blk                      # the ctypes array (base of every view) keeps the block alive
    return np.frombuffer(buf, dtype=np.float64, count=n).reshape(shape)


def hptr(a):
    """Host pointer of a C-contiguous float64/int32 numpy array (or None)."""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return ctypes.c_void_p(a.ctypes.data)


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


_SLOT_ARRAYS = {}


def slot_array(par_names):
    key = tuple(par_names)
    a = _SLOT_ARRAYS.get(key)
    if a is None:
        try:
            a = i32([SLOT[p] for p in key])
        except KeyError as e:
            raise ValueError("unknown parameter name %s (known: %s)" % (e, sorted(SLOT)))
        a.setflags(write=False)
        _SLOT_ARRAYS[key] = a
    return a


def stream_ptr(stream=None):
    """cudaStream_t of a torch stream (default: torch's current stream)."""
    import torch
    s = torch.cuda.current_stream() if stream is None else stream
    return ctypes.c_void_p(s.cuda_stream)


def dptr(t):
    """Device pointer of a contiguous float64 CUDA tensor (or None)."""
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "need a contiguous CUDA tensor"
    return ctypes.c_void_p(t.data_ptr())


class Plan:
    """Owns an ngrtd_plan: input series + lag tables of one (model pair, tracer set) resident in HBM."""

    def __init__(self, series, tracers, mod_type1, mod_type2=False, lag_index=None, dtp=0.0, device=-1):
        series = f64(series)
        if series.ndim == 1:
            series = series.reshape(-1, 1)
        self.L, self.nseries = series.shape
        for m in (mod_type1, mod_type2):
            if m and m not in MOD:
                raise ValueError("unknown mod_type %r (known: %s)" % (m, sorted(MOD)))
        if not mod_type1:
            raise ValueError("mod_type1 is required")
        arr = (Tracer * len(tracers))()
        for i, t in enumerate(tracers):
            ra = t.get("rad_accum", False)
            if ra not in ACC:
                raise ValueError("unknown rad_accum %r" % (ra,))
            arr[i] = Tracer(int(t.get("series", -1)), ACC[ra], float(t.get("lam", 0.0)),
                            int(bool(t.get("use_thalf_cfc", False))), int(bool(t.get("use_lamsf6", False))))
        li = None if lag_index is None else f64(lag_index)
        h = ctypes.c_void_p()
        check(lib.ngrtd_plan_create(ctypes.byref(h), self.L, self.nseries, hptr(series), hptr(li), float(dtp),
                                    len(tracers), arr, MOD[mod_type1], MOD[mod_type2] if mod_type2 else 0, device))
        self.handle = h
        self.ntracer = len(tracers)
        self._inflight = {}

    def close(self):
        if getattr(self, "handle", None):
            lib.ngrtd_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- host buffers (numpy in / numpy out) ----
    def forward_host(self, theta, par_names):
        theta = f64(np.atleast_2d(theta))
        B, ndim = theta.shape
        assert ndim == len(par_names)
        out = np.empty((B, self.ntracer))
        check(lib.ngrtd_forward_host(self.handle, hptr(theta), B, ndim, hptr(slot_array(par_names)), hptr(out)))
        return out

    def forward_loglik_host(self, theta, par_names, obs_mu, obs_sd, kind="normal", nu=None, want_model=False,
                            logp_out=None):
        theta = f64(np.atleast_2d(theta))
        B, ndim = theta.shape
        logp = np.empty(B) if logp_out is None else logp_out
        model = np.empty((B, self.ntracer)) if want_model else None
        nu_a = None if nu is None else f64(np.broadcast_to(nu, (B,)))
        check(lib.ngrtd_forward_loglik_host(self.handle, hptr(theta), B, ndim, hptr(slot_array(par_names)), LIK[kind],
                                            hptr(f64(obs_mu)), hptr(f64(obs_sd)), hptr(nu_a), hptr(logp), hptr(model)))
        return (logp, model) if want_model else logp

    HOST_SLOTS = 4

    def forward_loglik_host_submit(self, theta, par_names, obs_mu, obs_sd, kind="normal", nu=None, want_model=False,
                                   logp_out=None, model_out=None, slot=0):
        """Asynchronous form of forward_loglik_host for independent batches (Monte-Carlo sweeps, posterior-predictive
        loops): enqueue copy-in -> kernel -> copy-out for this batch and return at once; up to HOST_SLOTS batches are in
        flight and their copies overlap the kernels of their neighbours.  Returns (logp[, model]) arrays that are valid
        after host_wait(slot).  Pass pinned (page-locked) arrays for theta / logp_out / model_out to get the overlap."""
        theta = f64(np.atleast_2d(theta))
        B, ndim = theta.shape
        logp = np.empty(B) if logp_out is None else logp_out
        model = (np.empty((B, self.ntracer)) if model_out is None else model_out) if (want_model or model_out is not None) else None
        nu_a = None if nu is None else f64(np.broadcast_to(nu, (B,)))
        check(lib.ngrtd_forward_loglik_host_submit(self.handle, hptr(theta), B, ndim, hptr(slot_array(par_names)), LIK[kind],
                                                   hptr(f64(obs_mu)), hptr(f64(obs_sd)), hptr(nu_a), hptr(logp), hptr(model),
                                                   int(slot)))
        self._inflight[int(slot)] = (theta, nu_a, logp, model)       # keep the buffers alive until host_wait
        return (logp, model) if model is not None else logp

    def host_wait(self, slot=0):
        check(lib.ngrtd_host_wait(self.handle, int(slot)))
        self._inflight.pop(int(slot), None)

    # ---- device buffers (torch CUDA tensors in / out, no synchronisation) ----
    def forward_dev(self, theta_t, par_names, out_t=None, stream=None):
        import torch
        B, ndim = theta_t.shape
        if out_t is None:
            out_t = torch.empty((B, self.ntracer), dtype=torch.float64, device=theta_t.device)
        check(lib.ngrtd_forward_dev(self.handle, dptr(theta_t), B, ndim, hptr(slot_array(par_names)), dptr(out_t),
                                    stream_ptr(stream)))
        return out_t

    def forward_loglik_dev(self, theta_t, par_names, obs_mu, obs_sd, kind="normal", nu_t=None, logp_t=None,
                           model_t=None, stream=None):
        import torch
        B, ndim = theta_t.shape
        if logp_t is None:
            logp_t = torch.empty((B,), dtype=torch.float64, device=theta_t.device)
        check(lib.ngrtd_forward_loglik_dev(self.handle, dptr(theta_t), B, ndim, hptr(slot_array(par_names)), LIK[kind],
                                           hptr(f64(obs_mu)), hptr(f64(obs_sd)), dptr(nu_t), dptr(logp_t),
                                           dptr(model_t), stream_ptr(stream)))
        return logp_t
