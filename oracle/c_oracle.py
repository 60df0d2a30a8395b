"""TEST INFRASTRUCTURE ONLY -- ctypes loader of the plain-C oracle (oracle/ngrtd_oracle.c).

Used by tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs (as the checker and
as the timed CPU port).  Never imported by the product package.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "liboracle.so")
MOD = {False: 0, None: 0, "piston": 1, "exponential": 2, "exp_pist_flow": 3, "dispersion": 4}
SLOT = {"tau1": 0, "tau2": 1, "f1": 2, "f2": 3, "eta1": 4, "eta2": 5, "D1": 6, "D2": 7, "J": 8, "thalf_cfc": 9, "lamsf6": 10}
ACC = {False: 0, None: 0, "3He": 1, "4He": 2}
GAS = {"He": 0, "Ne": 1, "Ar": 2, "Kr": 3, "Xe": 4}


class OTracer(ctypes.Structure):
    _fields_ = [("series", ctypes.c_int32), ("rad_accum", ctypes.c_int32), ("lam", ctypes.c_double),
                ("use_thalf_cfc", ctypes.c_int32), ("use_lamsf6", ctypes.c_int32)]


def build(force=False):
    src = os.path.join(_HERE, "ngrtd_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "_build/liboracle.so"], stdout=subprocess.DEVNULL)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(ctypes.c_void_p)


def forward(series, tracers, mod1, mod2, theta, par_names, lag_index=None, dtp=0.0, nthreads=0):
    """series [L, nseries] newest-first; tracers: list of dicts (series, rad_accum, lam, use_thalf_cfc, use_lamsf6)."""
    series = np.ascontiguousarray(series, dtype=np.float64)
    if series.ndim == 1:
        series = series.reshape(-1, 1)
    L, ns = series.shape
    theta = np.ascontiguousarray(np.atleast_2d(theta), dtype=np.float64)
    B, ndim = theta.shape
    arr = (OTracer * len(tracers))()
    for i, t in enumerate(tracers):
        arr[i] = OTracer(int(t.get("series", -1)), ACC[t.get("rad_accum", False)], float(t.get("lam", 0.0)),
                         int(bool(t.get("use_thalf_cfc", False))), int(bool(t.get("use_lamsf6", False))))
    slots = np.ascontiguousarray([SLOT[p] for p in par_names], dtype=np.int32)
    li = None if lag_index is None else np.ascontiguousarray(lag_index, dtype=np.float64)
    out = np.empty((B, len(tracers)))
    f = lib().oracle_forward
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_double, ctypes.c_int32,
                  ctypes.POINTER(OTracer), ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int32,
                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int32]
    f(L, ns, _p(series), _p(li), float(dtp), len(tracers), arr, MOD[mod1], MOD[mod2], _p(theta), B, ndim, _p(slots),
      _p(out), int(nthreads))
    return out


def ce(what, gases, E, T, Ae, F, P=None, S=0.0):
    E, T, Ae, F = [np.ascontiguousarray(np.atleast_1d(v), dtype=np.float64) for v in (E, T, Ae, F)]
    B = T.shape[0]
    g = np.ascontiguousarray([GAS[x] for x in gases], dtype=np.int32)
    Pa = None if P is None else np.ascontiguousarray(P, dtype=np.float64)
    out = np.empty((B, len(gases)))
    f = lib().oracle_ce
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_int32, ctypes.c_int32] + [ctypes.c_void_p] * 6 + [ctypes.c_double, ctypes.c_int64, ctypes.c_void_p]
    f(int(what), len(gases), _p(g), _p(E), _p(T), _p(Ae), _p(F), _p(Pa), float(S), B, _p(out))
    return out


def loglik(kind, mu, obs, sd, nu=None):
    mu = np.ascontiguousarray(np.atleast_2d(mu), dtype=np.float64)
    B, T = mu.shape
    obs = np.ascontiguousarray(obs, dtype=np.float64)
    sd = np.ascontiguousarray(sd, dtype=np.float64)
    nu_a = None if nu is None else np.ascontiguousarray(np.broadcast_to(nu, (B,)), dtype=np.float64)
    out = np.empty(B)
    f = lib().oracle_loglik
    f.restype = ctypes.c_int
    f.argtypes = [ctypes.c_int32, ctypes.c_int32] + [ctypes.c_void_p] * 4 + [ctypes.c_int64, ctypes.c_void_p]
    f({"normal": 0, "studentt": 1}[kind], T, _p(mu), _p(obs), _p(sd), _p(nu_a), B, _p(out))
    return out


def max_threads():
    f = lib().oracle_max_threads
    f.restype = ctypes.c_int
    return int(f())
