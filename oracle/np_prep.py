"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the reference's observation-ensemble propagation
(age_modeling_mcmc.prep.py:107-141, 242-447), member by member like the reference's own loops, on top of the CE / CFC
restatements of np_oracle.py.  Pinned against tests/golden/ens_dict_small.npz (the reference's lines executed verbatim by
oracle/gen_golden_r2.py).  Checker of noblegas_rtd_mcmc_b200/prep.py; never imported by the product."""
import numpy as np

import np_oracle as O

PARS = ['m', 'b', 'Ae', 'F', 'E', 'T']


def he_comps(obs, Ae, F, E, T, Rterr):
    """utils/noble_gas_utils.py:394-422 (ng_parse.He_comps) for one parameter set."""
    P = O.lapse_rate(np.array([E]))
    eq = O.equil_conc(["He"], np.array([T]), P)[0, 0]
    atm = O.ce_exc(["He"], E, T, Ae, F, True)[0, 0]
    ter = obs['He4'] - atm
    dele = 100 * (ter / atm)
    Ratm = 1.384e-6
    CF = 4.021e14 / (1 - 0.0)
    he3_trit = obs['He3'] - (obs['He4'] - ter) * Ratm + eq * Ratm * (1 - 0.983) - ter * Rterr
    return {'He4_eq': eq, 'He4_atm': atm, 'He4_ter': ter, 'He4_del': dele, 'He3_tu': he3_trit * CF}


def propagate(draws, field_obs, err, seed=10, n_members=50000, obs_list=('PLM1', 'PLM7', 'PLM6')):
    ix = {p: i for i, p in enumerate(PARS)}
    sel = [ix[p] for p in ('Ae', 'F', 'E', 'T')]
    np.random.seed(seed)
    par_map, par_ens = {}, {}
    for w in obs_list:
        p = np.array(draws[w], dtype=np.float64, copy=True)
        par_map[w] = [p[:, i].mean() for i in range(6)]
        np.random.shuffle(p)
        par_ens[w] = p[:n_members]
    ens = {k: {} for k in ('CFC11', 'CFC12', 'CFC113', 'SF6', 'He4_ter', 'He4_ter_del', 'H3_He3', 'He3', 'H3', 'H3_init')}
    mp = {k: {} for k in ('CFC11', 'CFC12', 'CFC113', 'SF6', 'He4_ter', 'H3_He3', 'He3')}
    for w in ('PLM1', 'PLM6', 'PLM7'):
        cm = np.array([field_obs[w][k] for k in ('CFC11', 'CFC12', 'CFC113')])
        Ae, F, E, T = np.array(par_map[w])[sel]
        m = O.cfc_corr('air', [11, 12, 113], E, T, Ae, F, cm[None, :])[0]
        for i, k in enumerate(('CFC11', 'CFC12', 'CFC113')):
            mp[k][w] = m[i]
        n = len(par_ens[w])
        unc = np.array([np.random.normal(0.0, cm[i] * err['CFC'], n) for i in range(3)]).T
        rows = []
        for i in range(n):
            Ae, F, E, T = par_ens[w][i, sel]
            rows.append(O.cfc_corr('air', [11, 12, 113], E, T, Ae, F, (cm + unc[i])[None, :])[0])
        rows = np.array(rows)
        for i, k in enumerate(('CFC11', 'CFC12', 'CFC113')):
            ens[k][w] = rows[:, i]
    for w in ('PLM1', 'PLM6', 'PLM7'):
        cm = field_obs[w]['SF6']
        Ae, F, E, T = np.array(par_map[w])[sel]
        mp['SF6'][w] = O.cfc_corr('air', [6], E, T, Ae, F, np.array([[cm]]))[0, 0]
        unc = np.random.normal(0.0, cm * err['SF6'], len(par_ens[w]))
        ens['SF6'][w] = np.array([O.cfc_corr('air', [6], *par_ens[w][i, sel][[2, 3, 0, 1]], np.array([[cm + unc[i]]]))[0, 0]
                                  for i in range(len(par_ens[w]))])
    low_, high_ = np.log10(2.e-8) - 1.0, np.log10(2.e-8) + 1.0
    beta_ = np.random.beta(2, 2, len(par_ens['PLM7']))
    Rrv = 10 ** (beta_ * (high_ - low_) + low_)
    Rterr = 10 ** (np.log10(Rrv).mean())
    he_map = {}
    for w in obs_list:
        Ae, F, E, T = np.array(par_map[w])[sel]
        he_map[w] = he_comps(field_obs[w], Ae, F, E, T, Rterr)
        mp['He4_ter'][w] = he_map[w]['He4_ter']
        mp['He3'][w] = he_map[w]['He3_tu']

    def sample(Rl):
        out = {}
        for w in obs_list:
            rows = []
            for i in range(len(par_ens[w])):
                o = dict(He4=field_obs[w]['He4'], He3=field_obs[w]['He3'])
                o['He4'] += np.random.normal(0, field_obs[w]['He4'] * err['He4'])
                o['He3'] += np.random.normal(0, field_obs[w]['He3'] * err['He3'])
                Ae, F, E, T = par_ens[w][i, sel]
                rows.append(he_comps(o, Ae, F, E, T, Rl[i]))
            out[w] = {k: np.array([r[k] for r in rows]) for k in rows[0]}
        return out
    cr = sample(np.ones_like(Rrv) * Rterr)
    rv = sample(Rrv)
    marg = {}
    for rl in np.concatenate(([-12], np.linspace(np.log10(Rrv.min()), np.log10(Rrv.max()), 5))):
        marg[-rl] = {w: v['He3_tu'] for w, v in sample(np.ones_like(Rrv) * 10 ** rl).items()}
    for w in obs_list:
        mp['H3_He3'][w] = field_obs[w]['H3'] / he_map[w]['He3_tu']
        h3 = np.random.normal(field_obs[w]['H3'], field_obs[w]['H3'] * err['H3'], len(rv[w]['He3_tu']))
        ens['H3'][w] = h3
        ens['H3_He3'][w] = h3 / rv[w]['He3_tu']
        ens['H3_init'][w] = rv[w]['He3_tu'] + h3
        ens['He4_ter'][w] = rv[w]['He4_ter']
        ens['He4_ter_del'][w] = rv[w]['He4_del']
        ens['He3'][w] = rv[w]['He3_tu']
    return mp, ens, marg, Rterr
