"""Drop-in for the reference's utils/cfc_utils.py (`cfc_ce_corr`, `sf6_ce_corr`): CFC-11/12/113 and SF6 solubility and
closed-system excess-air corrections on the B200 (libngrtd.so: ngrtd_cfc_*).  Scalar E/T/Ae/F give the reference's
results (ndarray over species for the CFC class, float for SF6); array-valued E/T/Ae/F of shape [B] -- e.g. the 50,000
noble-gas posterior draws of age_modeling_mcmc.prep.py:242-303 -- give [B, nspecies] / [B] in one launch."""
import numpy as np

from . import _lib

_W = {"air": 0, "aq": 1, "exc": 2, "K": 3}


def _run(what, species, E, T, Ae, F, S, X):
    B = 1
    scalar = True
    for a in (E, T, Ae, F):
        if np.ndim(a) > 0:
            scalar = False
            B = max(B, np.size(a))
    ns = len(species)
    cols = [_lib.f64(np.broadcast_to(np.asarray(a, dtype=np.float64), (B,))) for a in (E, T, Ae, F)]
    Xa = None
    if X is not None:
        Xa = np.asarray(X, dtype=np.float64)
        if Xa.ndim == 2 and Xa.shape[0] == B and B > 1:
            scalar = False
        Xa = _lib.f64(np.broadcast_to(Xa, (B, ns)) if Xa.ndim < 2 or Xa.shape != (B, ns) else Xa)
    out = np.empty((B, ns))
    sp = _lib.i32(species)
    _lib.check(_lib.lib.ngrtd_cfc_host(_W[what], ns, _lib.hptr(sp), _lib.hptr(cols[0]), _lib.hptr(cols[1]), _lib.hptr(cols[2]),
                                       _lib.hptr(cols[3]), _lib.hptr(Xa), float(S), B, _lib.hptr(out)))
    return out, scalar


class cfc_ce_corr():
    def __init__(self, cfc_num, E, T, Ae, F, S=0.0):
        for c in cfc_num:
            if c not in (11, 12, 113):
                raise ValueError("cfc_num entries must be 11, 12 or 113")
        self.cfc = list(cfc_num)
        self.E = E
        self.T = T
        self.Ae = np.asarray(Ae) * 1000. if np.ndim(Ae) else Ae * 1000.     # ccSTP/g -> ccSTP/kg (cfc_utils.py:29)
        self.F = F
        self.S = S
        self.P = None

    def _call(self, what, X=None):
        Ae_g = np.asarray(self.Ae) / 1000. if np.ndim(self.Ae) else self.Ae / 1000.
        out, scalar = _run(what, self.cfc, self.E, self.T, Ae_g, self.F, self.S, X)
        return out[0] if scalar else out

    def vapor_pressure_atm(self):                        # :35-52: the CE kernel's Antoine pressure [GPa] in atmospheres
        from .noble_gas_utils import _ce_pressure
        return _ce_pressure("P_vapor", self.T) / 0.000101325

    def lapse_rate_atm(self):                            # :54-60: the CE kernel's lapse-rate pressure [GPa] in atmospheres
        from .noble_gas_utils import _ce_pressure
        self.P = _ce_pressure("P_lapse", self.E) / 0.000101325
        return self.P

    def solubility_cfc(self):                            # :62-83
        return self._call("K")

    def equil_air_conc_cfc(self, C_meas):                # :85-105
        return self._call("air", C_meas)

    def equil_aq_conc_cfc(self, z_i):                    # :107-124
        return self._call("aq", z_i)

    def ce_exc_conc_cfc(self, z_i):                      # :126-143
        return self._call("exc", z_i)

    def update_pars(self, T, E, Ae, F):                  # :145-151
        self.T = T
        self.E = E
        self.Ae = np.asarray(Ae) * 1000 if np.ndim(Ae) else Ae * 1000
        self.F = F
        self.vapor_pressure_atm()
        self.lapse_rate_atm()


class sf6_ce_corr():
    def __init__(self, E, T, Ae, F, S=0.0):
        self.E = E
        self.T = T
        self.Ae = np.asarray(Ae) * 1000. if np.ndim(Ae) else Ae * 1000.     # :164
        self.F = F
        self.S = S
        self.P = None

    def _call(self, what, X=None):
        Ae_g = np.asarray(self.Ae) / 1000. if np.ndim(self.Ae) else self.Ae / 1000.
        Xa = None if X is None else (np.asarray(X, dtype=np.float64).reshape(-1, 1) if np.ndim(X) else X)
        out, scalar = _run(what, [6], self.E, self.T, Ae_g, self.F, self.S, Xa)
        return float(out[0, 0]) if scalar and np.ndim(X) == 0 else out[:, 0]

    vapor_pressure_atm = cfc_ce_corr.vapor_pressure_atm
    lapse_rate_atm = cfc_ce_corr.lapse_rate_atm

    def solubility_sf6(self):                            # :196-209
        return self._call("K")

    def equil_air_conc_sf6(self, C_meas):                # :228-247
        return self._call("air", C_meas)

    def equil_aq_conc_sf6(self, z_i):                    # :262-278
        return self._call("aq", z_i)

    def ce_exc_conc_sf6(self, z_i):                      # :280-296
        return self._call("exc", z_i)

    update_pars = cfc_ce_corr.update_pars
