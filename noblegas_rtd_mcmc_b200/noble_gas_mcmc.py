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
