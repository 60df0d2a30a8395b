"""Drop-in for the hot-path part of the reference's utils/noble_gas_utils.py: `atm_std`, `noble_gas_fun`
(closed-equilibrium model, :75-261) and `J_flux` (:335-348), evaluated on the B200 through libngrtd.so.

Scalar arguments give scalar results exactly as in the reference; array-valued E/T/Ae/F of shape [B] give
ndarray results of shape [B] per gas (additive behaviour).
"""
import numpy as np

from . import _lib

# Dry air mixing ratios, Porcelli et al. 2002 (utils/noble_gas_utils.py:38-73 of the reference)
atm_std = {'N2': 0.781, 'O2': 0.209, 'Ar': 9.34e-3, 'CO2': 3.7e-4, 'Ne': 1.818e-5, 'He': 5.24e-6, 'CH4': 1.5e-6,
           'Kr': 1.14e-6, 'H2': 7e-7, 'N2O': 3e-7, 'CO': 1e-7, 'Xe': 8.7e-8, 'Rn': 6e-20,
           'He3': 5.24e-6 * 0.000140 / 100, 'He4': 5.24e-6,
           'Ne20': 1.818e-5 * 90.50 / 100, 'Ne21': 1.818e-5 * 0.268 / 100, 'Ne22': 1.818e-5 * 9.23 / 100,
           'Ar36': 9.34e-3 * 0.3364 / 100, 'Ar38': 9.34e-3 * 0.0632 / 100, 'Ar40': 9.34e-3 * 99.60 / 100,
           'Kr78': 1.14e-6 * 0.3469 / 100, 'Kr80': 1.14e-6 * 2.2571 / 100, 'Kr82': 1.14e-6 * 11.523 / 100,
           'Kr83': 1.14e-6 * 11.477 / 100, 'Kr84': 1.14e-6 * 57.00 / 100, 'Kr86': 1.14e-6 * 17.398 / 100,
           'Xe124': 8.7e-8 * 0.0951 / 100, 'Xe126': 8.7e-8 * 0.0887 / 100, 'Xe128': 8.7e-8 * 1.919 / 100,
           'Xe129': 8.7e-8 * 26.44 / 100, 'Xe130': 8.7e-8 * 4.070 / 100, 'Xe131': 8.7e-8 * 21.22 / 100,
           'Xe132': 8.7e-8 * 26.89 / 100, 'Xe134': 8.7e-8 * 10.430 / 100, 'Xe136': 8.7e-8 * 8.857 / 100}

_WHAT = {"ce_true": 0, "ce_false": 1, "eq_dry": 2, "eq_wet": 3, "K": 4, "P_lapse": 5, "P_vapor": 6}


def _gas_id(gas):
    g = gas[0:2]                                   # the reference keys solubility on the first two letters (:135)
    if g not in _lib.GAS:
        raise ValueError("unknown noble gas %r (known: He, Ne, Ar, Kr, Xe)" % (gas,))
    return _lib.GAS[g]


def _ce_pressure(what, x):
    """lapse_rate() / vapor_pressure() of the reference (:103-113, :184-199) [GPa], evaluated by the CE kernel
    (selectors 5 / 6 of ngrtd_ce_host); x = E [m] or T [C], scalar or array."""
    v = np.atleast_1d(np.asarray(x, dtype=np.float64))
    out = np.empty((v.size, 1))
    zeros = np.zeros(v.size)
    E, T = (v.ravel(), zeros) if what == "P_lapse" else (None, v.ravel())
    _lib.check(_lib.lib.ngrtd_ce_host(_WHAT[what], 1, _lib.hptr(_lib.i32([1])), _lib.hptr(_lib.f64(E)) if E is not None else None,
                                      _lib.hptr(_lib.f64(T)), None, None, None, 0.0, v.size, _lib.hptr(out)))
    return float(out[0, 0]) if np.ndim(x) == 0 else out[:, 0].reshape(np.shape(x))


class noble_gas_fun():
    def __init__(self, gases, E, T, Ae, F, P, S=0.0):
        self.gases = gases
        self.E = E
        self.T = T
        self.Ae = Ae
        self.F = F
        self.S = S
        self.P = self.parse_P(P)
        self.obs_dict = None
        self.err_dict = None

    def parse_P(self, P):                          # :91-100
        if isinstance(P, str):
            if P == '1atm':
                return 0.000101325
            if P == 'lapse_rate':
                return self.lapse_rate()
            raise ValueError("P must be '1atm', 'lapse_rate' or a pressure in GPa")
        return P

    def lapse_rate(self):                          # :103-113, evaluated by the CE kernel (selector 5, P = NULL)
        return _ce_pressure("P_lapse", self.E)

    def _run(self, what, gases):
        ids = _lib.i32([_gas_id(g) for g in gases])
        args = [self.E, self.T, self.Ae, self.F, self.P]
        B = 1
        scalar = True
        for a in args:
            if np.ndim(a) > 0:
                scalar = False
                B = max(B, np.size(a))
        E, T, Ae, F, P = [_lib.f64(np.broadcast_to(np.asarray(0.0 if a is None else a, dtype=np.float64), (B,))) for a in args]
        out = np.empty((B, len(gases)))
        _lib.check(_lib.lib.ngrtd_ce_host(_WHAT[what], len(gases), _lib.hptr(ids), _lib.hptr(E), _lib.hptr(T), _lib.hptr(Ae),
                                          _lib.hptr(F), _lib.hptr(P), float(self.S), B, _lib.hptr(out)))
        return out, scalar

    def _as_dict(self, what, gases):
        out, scalar = self._run(what, gases)
        return {g: (float(out[0, i]) if scalar else out[:, i].copy()) for i, g in enumerate(gases)}

    def solubility(self, gas):                     # :117-180
        out, scalar = self._run("K", [gas])
        return float(out[0, 0]) if scalar else out[:, 0]

    def vapor_pressure(self):                      # :184-199, evaluated by the CE kernel (selector 6)
        return _ce_pressure("P_vapor", self.T)

    def equil_conc(self):                          # :202-213
        return self._as_dict("eq_wet", list(self.gases))

    def equil_conc_dry(self):                      # :216-232
        return self._as_dict("eq_dry", list(self.gases))

    def ce_exc(self, add_eq_conc):                 # :235-253
        return self._as_dict("ce_true" if add_eq_conc else "ce_false", list(self.gases))

    def update_pars(self, T, Ae, F, E):            # :256-261
        self.T = T
        self.Ae = Ae
        self.F = F
        self.E = E
        self.P = self.lapse_rate()


def J_flux(Del, rho_r, rho_w, U, Th, phi):
    """He accumulation rate (ccSTP/g_water/yr), Porcelli 2002 p. 648 (utils/noble_gas_utils.py:335-348)."""
    PU = 1.19e-13
    PTh = 2.88e-14
    return Del * rho_r / rho_w * (U * PU + Th * PTh) * ((1 - phi) / phi)


class ng_parse():
    """Helium component separation of the reference (utils/noble_gas_utils.py:354-422, `He_comps`), batched: array-valued
    Ae/F/E/T (e.g. the 50,000 posterior draws the reference loops over at age_modeling_mcmc.prep.py:356-385) give
    array-valued components in `obs_dict_`.  One CE launch replaces three `noble_gas_fun` evaluations."""

    def __init__(self, obs_dict, Ae, F, E, T):
        self.obs_dict = obs_dict          # keys He4, He3 (single well)
        self.Ae = Ae
        self.F = F
        self.E = E
        self.T = T
        self.obs_dict_ = obs_dict.copy()

    def He_comps(self, Rterr):
        He4_obs = self.obs_dict['He4']
        He3_obs = self.obs_dict['He3']
        ng_ = noble_gas_fun(gases=['He'], E=self.E, T=self.T, Ae=self.Ae, F=self.F, P='lapse_rate')
        He4_eq = ng_.equil_conc()['He']                    # atmospheric equilibrium                      (:400)
        He4_ex = ng_.ce_exc(add_eq_conc=False)['He']       # excess-air component                         (:401)
        He4_atm = ng_.ce_exc(add_eq_conc=True)['He']       # equilibrium + excess air                     (:402)
        He4_ter = He4_obs - He4_atm                        # terrigenic 4He                               (:403)
        He4_del = 100 * (He4_ter / He4_atm)                # (:406)
        self.obs_dict_['He4_eq'] = He4_eq
        self.obs_dict_['He4_atm'] = He4_atm
        self.obs_dict_['He4_ter'] = He4_ter
        self.obs_dict_['He4_del'] = He4_del
        Ratm = 1.384e-6                                    # 3He/4He in the atmosphere                    (:413)
        S = 0.0
        CF = 4.021e14 / (1 - S)
        He3_trit = He3_obs - (He4_obs - He4_ter) * Ratm + He4_eq * Ratm * (1 - 0.983) - He4_ter * Rterr   # (:420)
        self.obs_dict_['He3_tu'] = He3_trit * CF
        self.He4_ex = He4_ex
