"""TEST INFRASTRUCTURE ONLY -- round-2 golden vectors from the UNTOUCHED reference (/root/reference), in addition to
oracle/gen_golden.py (whose fixtures and random stream are left exactly as they are).

    python oracle/gen_golden_r2.py

Writes
  noblegas_rtd_mcmc_b200/data/field_obs_plm.json   the field observations prep.py reads from Field_Data/*.xlsx (:60-76), parsed
                                                    from the xlsx XML (openpyxl is not installed): data of the reference
  tests/golden/ce_salinity.npz                      noble_gas_fun / cfc_ce_corr / sf6_ce_corr with S != 0 (Setchenow terms)
  tests/golden/ens_dict_small.npz                   the reference's observation-ensemble propagation
                                                    (age_modeling_mcmc.prep.py:242-489, executed VERBATIM from the reference
                                                    file -- nothing is copied into this repo) on synthetic CE posteriors of 64
                                                    draws per well: inputs, the resulting ens_dict / map_dict values.
The reference cannot travel to the GPU box, so the vectors are committed as small fixtures.
"""
import copy
import json
import os
import re
import sys
import tempfile
import zipfile

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import ref_shims  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
DATA = os.path.join(ROOT, "noblegas_rtd_mcmc_b200", "data")
WELLS = ["PLM1", "PLM7", "PLM6"]                 # obs_list of prep.py:81
PARS = ["m", "b", "Ae", "F", "E", "T"]           # prep.py:117


def read_xlsx(path, header_row):
    """rows of sheet1 as dicts keyed by the header row (0-based index of the row that holds the column names)"""
    z = zipfile.ZipFile(path)
    ss = z.read("xl/sharedStrings.xml").decode()
    strings = [re.sub(r"<[^>]+>", "", m) for m in re.findall(r"<si>(.*?)</si>", ss, flags=re.S)]
    sh = z.read("xl/worksheets/sheet1.xml").decode()
    rows = []
    for r in re.findall(r"<row[^>]*>(.*?)</row>", sh, flags=re.S):
        cells = {}
        for col, attr, body in re.findall(r'<c r="([A-Z]+)\d+"([^>]*?)(?:/>|>(.*?)</c>)', r, flags=re.S):
            v = re.search(r"<v>(.*?)</v>", body or "")
            if v is None:
                continue
            val = v.group(1)
            cells[col] = strings[int(val)] if 't="s"' in attr else val
        rows.append(cells)
    hdr = rows[header_row]
    out = []
    for r in rows[header_row + 1:]:
        if "A" in r:
            out.append({hdr[c]: r[c] for c in r if c in hdr})
    return out


def field_obs():
    fd = os.path.join(ref_shims.REFERENCE_ROOT, "Field_Data")
    tr = {r["Sample"]: r for r in read_xlsx(os.path.join(fd, "PLM_tracers_2021.xlsx"), 0) if r.get("Sample", "").startswith("PLM")}
    ng = {r["SiteID"]: r for r in read_xlsx(os.path.join(fd, "PLM_noblegas_2021.xlsx"), 1) if r.get("SiteID", "").startswith("PLM")}
    obs = {"source": "Field_Data/PLM_tracers_2021.xlsx and PLM_noblegas_2021.xlsx of the reference (values as stored; prep.py adds "
                     "1e-10 to the CFC / SF6 / 3H observations, :61-69)",
           "units": {"CFC": "pmol/kg", "SF6": "fmol/kg", "H3": "TU", "He": "ccSTP/g"}, "wells": {}}
    for w in sorted(tr):
        d = {k: float(tr[w][k]) for k in ("CFC11", "CFC12", "CFC113", "SF6", "H3")}
        if w in ng and "4He" in ng[w] and "3He" in ng[w]:
            d["He4"], d["He3"] = float(ng[w]["4He"]), float(ng[w]["3He"])
        obs["wells"][w] = d
    return obs


def main():
    ref_shims.install()
    import cfc_utils
    import noble_gas_utils as ng_utils
    os.makedirs(DATA, exist_ok=True)
    obs = field_obs()
    with open(os.path.join(DATA, "field_obs_plm.json"), "w") as f:
        json.dump(obs, f, indent=1)

    # ---- salinity goldens ----
    rng = np.random.default_rng(20261019)
    gases = ["He", "Ne", "Ar", "Kr", "Xe"]
    n = 60
    E = rng.uniform(0.0, 3300.0, n); T = rng.uniform(0.1, 30.0, n); Ae = 10 ** rng.uniform(-4, -1, n); F = 10 ** rng.uniform(-1, 0.5, n)
    S = np.where(np.arange(n) % 2 == 0, 35.0, 5.0)
    T[:3] = [64.9, 65.0, 70.0]                          # the Setchenow factor applies below 65 C only (:145)
    res = {k: np.empty((n, 5)) for k in ("ce_true", "eq_dry", "eq_wet", "K")}
    Cm = rng.uniform(0.2, 5.0, (n, 3)); Cs = rng.uniform(0.1, 3.0, n)
    cfc_air = np.empty((n, 3)); cfc_K = np.empty((n, 3)); sf6_air = np.empty(n); sf6_K = np.empty(n)
    with np.errstate(all="ignore"):
        for i in range(n):
            o = ng_utils.noble_gas_fun(gases=gases, E=E[i], T=T[i], Ae=Ae[i], F=F[i], P="lapse_rate", S=S[i])
            a, c, d = o.ce_exc(True), o.equil_conc_dry(), o.equil_conc()
            for j, g in enumerate(gases):
                res["ce_true"][i, j], res["eq_dry"][i, j], res["eq_wet"][i, j] = a[g], c[g], d[g]
                res["K"][i, j] = o.solubility(g)
            cc = cfc_utils.cfc_ce_corr(cfc_num=[11, 12, 113], E=E[i], T=min(T[i], 30.0), Ae=Ae[i], F=F[i], S=S[i])
            cfc_K[i] = cc.solubility_cfc(); cfc_air[i] = cc.equil_air_conc_cfc(Cm[i])
            s6 = cfc_utils.sf6_ce_corr(E=E[i], T=min(T[i], 30.0), Ae=Ae[i], F=F[i], S=S[i])
            sf6_K[i] = s6.solubility_sf6(); sf6_air[i] = s6.equil_air_conc_sf6(Cs[i])
    np.savez_compressed(os.path.join(GOLD, "ce_salinity.npz"), E=E, T=T, Ae=Ae, F=F, S=S, Cm=Cm, Cs=Cs, cfc_air=cfc_air, cfc_K=cfc_K,
                        sf6_air=sf6_air, sf6_K=sf6_K, **res)

    # ---- observation-ensemble propagation: prep.py:242-489 executed from the reference file ----
    N = 64
    prng = np.random.default_rng(7)
    draws = {}
    for w in WELLS:                                    # synthetic CE posteriors, same columns as prep.py:117 (m, b, Ae, F, E, T)
        draws[w] = np.stack([prng.normal(-146, 17, N), prng.uniform(2989, 3719, N), 10 ** prng.uniform(-3.2, -1.6, N),
                             10 ** prng.uniform(-0.9, 0.4, N), prng.uniform(2790, 3250, N), prng.uniform(0.5, 7.5, N)], axis=1)
    src = open(os.path.join(ref_shims.REFERENCE_ROOT, "age_modeling_mcmc.prep.py")).read().splitlines()
    block = "\n".join(src[241:489])                    # lines 242..489: cfc_wells = ... up to (not including) "# Save them"
    wells_all = obs["wells"]
    idx = [w for w in wells_all]
    cfc_obs = pd.DataFrame({k: [wells_all[w][k] for w in idx] for k in ("CFC11", "CFC12", "CFC113")}, index=idx) + 1.e-10
    sf6_obs = pd.DataFrame({"SF6": [wells_all[w]["SF6"] for w in idx]}, index=idx) + 1.e-10
    h3_obs = pd.DataFrame({"H3": [wells_all[w]["H3"] for w in idx]}, index=idx) + 1.e-10
    he_idx = [w for w in idx if "He4" in wells_all[w]]
    he_obs = pd.DataFrame({"He4": [wells_all[w]["He4"] for w in he_idx], "He3": [wells_all[w]["He3"] for w in he_idx]}, index=he_idx)
    err_dict = {'CFC11': 0.05, 'CFC12': 0.05, 'CFC113': 0.05, 'CFC': 0.05, 'SF6': 0.05, 'H3': 0.08, 'He4': 0.02, 'R': 0.015, 'He3': 0.03}
    par_msk = dict(zip(PARS, np.arange(len(PARS))))
    np.random.seed(10)                                 # prep.py:110
    par_ens, par_map = {}, {}
    for w in WELLS:                                    # prep.py:123-137 without the arviz read
        p_ens = draws[w].copy()
        par_map[w] = [float(p_ens[:, i].mean()) for i in range(len(PARS))]
        np.random.shuffle(p_ens)
        p_ens = p_ens[:50000, :]
        par_ens[w] = np.array([p_ens[:, i] for i in range(len(PARS))]).T
    par_ce = ['Ae', 'F', 'E', 'T']
    ns = dict(np=np, pd=pd, copy=copy, cfc_utils=cfc_utils, ng_utils=ng_utils, cfc_obs=cfc_obs, sf6_obs=sf6_obs, h3_obs=h3_obs,
              he_obs=he_obs, err_dict=err_dict, par_map=par_map, par_ens=par_ens, par_msk=par_msk, par_ce=par_ce,
              p_inds=[par_msk[p] for p in par_ce], obs_list=WELLS)
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as tmp:
        os.chdir(tmp)
        try:
            with np.errstate(all="ignore"):
                exec(compile(block, "age_modeling_mcmc.prep.py[242:489]", "exec"), ns)
        finally:
            os.chdir(cwd)
    out = {"N": np.array(N), "wells": np.array(WELLS)}
    for w in WELLS:
        out["draws/" + w] = draws[w]
        out["par_map/" + w] = np.array(par_map[w])
        for key in ("CFC11", "CFC12", "CFC113", "SF6", "He4_ter", "He4_ter_del", "H3_He3", "He3", "H3", "H3_init"):
            out["ens/%s/%s" % (key, w)] = np.asarray(ns["ens_dict"][key][w]).ravel()
        for key in ("CFC11", "CFC12", "CFC113", "SF6", "He4_ter", "H3_He3", "He3"):
            out["map/%s/%s" % (key, w)] = np.asarray(ns["map_dict"][key].loc[w]).ravel()
        for rl, d in ns["he3_ens_marg"].items():
            out["marg/%.6f/%s" % (rl, w)] = np.asarray(d[w]).ravel()
    out["Rterr"] = np.array(ns["Rterr"])
    np.savez_compressed(os.path.join(GOLD, "ens_dict_small.npz"), **out)
    print("written:", os.path.join(DATA, "field_obs_plm.json"), "ce_salinity.npz", "ens_dict_small.npz")
    for k in ("ens/CFC12/PLM1", "ens/He4_ter/PLM6", "ens/H3_He3/PLM7", "map/SF6/PLM1"):
        print(" ", k, out[k][:3])


if __name__ == "__main__":
    main()
