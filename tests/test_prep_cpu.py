"""CPU: the numpy restatement of the reference's observation-ensemble propagation (oracle/np_prep.py) against the golden
vectors produced by the reference's own lines (tests/golden/ens_dict_small.npz, oracle/gen_golden_r2.py), the packaged
field observations, and the S != 0 (Setchenow) golden vectors against the CE / CFC restatements."""
import json
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _field_obs():
    d = json.load(open(os.path.join(ROOT, "noblegas_rtd_mcmc_b200", "data", "field_obs_plm.json")))["wells"]
    out = {}
    for w, v in d.items():
        o = {k: v[k] + 1.e-10 for k in ("CFC11", "CFC12", "CFC113", "SF6", "H3")}
        if "He4" in v:
            o["He4"], o["He3"] = v["He4"], v["He3"]
        out[w] = o
    return out


def test_oracle_prep_matches_the_reference_lines():
    import np_prep
    z = np.load(os.path.join(GOLD, "ens_dict_small.npz"))
    wells = [str(w) for w in z["wells"]]
    draws = {w: z["draws/" + w] for w in wells}
    err = {'CFC': 0.05, 'SF6': 0.05, 'H3': 0.08, 'He4': 0.02, 'He3': 0.03}
    mp, ens, marg, Rterr = np_prep.propagate(draws, _field_obs(), err)
    assert abs(Rterr - float(z["Rterr"])) <= 1e-14 * Rterr
    for k in z.files:
        parts = k.split("/")
        if parts[0] == "ens":
            got = ens[parts[1]][parts[2]]
        elif parts[0] == "map":
            got = np.atleast_1d(mp[parts[1]][parts[2]])
        elif parts[0] == "marg":
            key = min(marg, key=lambda r: abs(r - float(parts[1])))
            got = marg[key][parts[2]]
        else:
            continue
        want = z[k]
        assert got.shape == want.shape, k
        assert np.max(np.abs(got - want) / np.maximum(np.abs(want), 1e-300)) < 1e-11, k


def test_packaged_series_and_observations_match_the_fixtures():
    a = np.load(os.path.join(GOLD, "c_in_head.npz"))
    b = np.load(os.path.join(ROOT, "noblegas_rtd_mcmc_b200", "data", "c_in_head.npz"))
    assert sorted(a.files) == sorted(b.files)
    for k in a.files:
        assert np.array_equal(a[k], b[k]), k
    post = json.load(open(os.path.join(GOLD, "ng_posterior.json")))["wells"]
    pk = json.load(open(os.path.join(ROOT, "noblegas_rtd_mcmc_b200", "data", "ng_obs_plm.json")))["wells"]
    fo = json.load(open(os.path.join(ROOT, "noblegas_rtd_mcmc_b200", "data", "field_obs_plm.json")))["wells"]
    for w in ("PLM1", "PLM6", "PLM7"):
        assert pk[w]["obs"] == post[w]["obs"]
        assert fo[w]["He4"] == post[w]["obs"]["He"]          # the two xlsx sheets / the PANGA csv agree


def test_oracle_salinity_goldens():
    import np_oracle as O
    z = np.load(os.path.join(GOLD, "ce_salinity.npz"))
    gases = ["He", "Ne", "Ar", "Kr", "Xe"]
    for S in (5.0, 35.0):
        m = z["S"] == S
        E, T, Ae, F = z["E"][m], z["T"][m], z["Ae"][m], z["F"][m]
        K = np.stack([O.solubility(g, T, S) for g in gases], axis=1)
        assert np.max(np.abs(K / z["K"][m] - 1)) < 1e-13
        assert np.max(np.abs(O.ce_exc(gases, E, T, Ae, F, True, "lapse_rate", S) / z["ce_true"][m] - 1)) < 1e-12
        assert np.max(np.abs(O.equil_conc(gases, T, O.lapse_rate(E), S) / z["eq_wet"][m] - 1)) < 1e-13
        Tc = np.minimum(T, 30.0)
        assert np.max(np.abs(O.cfc_corr('K', [11, 12, 113], E, Tc, Ae, F, None, S) / z["cfc_K"][m] - 1)) < 1e-13
        assert np.max(np.abs(O.cfc_corr('air', [11, 12, 113], E, Tc, Ae, F, z["Cm"][m], S) / z["cfc_air"][m] - 1)) < 1e-12
        assert np.max(np.abs(O.cfc_corr('air', [6], E, Tc, Ae, F, z["Cs"][m][:, None], S)[:, 0] / z["sf6_air"][m] - 1)) < 1e-12
