"""Posterior diagnostics on host arrays posterior[chain, draw]: rank-normalised split R-hat, bulk ESS, MCSE, HDI
(Vehtari et al. 2021 -- the quantities ArviZ 0.11.4 `az.summary` reports for the reference at
age_ens_runs_mcmc/run_age_mcmc.py:234 and ng_interp/noble_gas_mcmc.py:450), plus the moment-based estimators used
when chains are too many to keep traces (per-chain Welford moments gathered across GPUs)."""
import numpy as np
from scipy import stats as _st


def _split(a):
    n = a.shape[1] // 2
    return np.concatenate([a[:, :n], a[:, a.shape[1] - n:]], axis=0)


def _z_scale(a):
    r = _st.rankdata(a, method="average").reshape(a.shape)
    return _st.norm.ppf((r - 0.375) / (a.size + 0.25))


def _autocov(x):
    n = x.shape[-1]
    m = 1 << int(np.ceil(np.log2(2 * n)))
    xc = x - x.mean(axis=-1, keepdims=True)
    f = np.fft.rfft(xc, m, axis=-1)
    ac = np.fft.irfft(f * np.conj(f), m, axis=-1)[..., :n]
    return ac / n


def _rhat_plain(a):
    m, n = a.shape
    cm = a.mean(axis=1)
    W = a.var(axis=1, ddof=1).mean()
    Bn = cm.var(ddof=1) if m > 1 else 0.0
    return float(np.sqrt(((n - 1) / n * W + Bn) / W))


def rhat(a):
    """Rank-normalised split R-hat (max of bulk and folded)."""
    a = np.asarray(a, dtype=np.float64)
    s = _split(a)
    bulk = _rhat_plain(_z_scale(s))
    folded = _rhat_plain(_z_scale(np.abs(s - np.median(s))))
    return max(bulk, folded)


def _ess_raw(a):
    m, n = a.shape
    if n < 4:
        return float("nan")
    acov = _autocov(a)
    cm = a.mean(axis=1)
    mean_var = acov[:, 0].mean() * n / (n - 1.0)
    var_plus = mean_var * (n - 1.0) / n
    if m > 1:
        var_plus += cm.var(ddof=1)
    rho = np.zeros(n)
    rho_even = 1.0
    rho[0] = rho_even
    rho_odd = 1.0 - (mean_var - acov[:, 1].mean()) / var_plus
    rho[1] = rho_odd
    t = 1
    while t < n - 3 and (rho_even + rho_odd) > 0.0:
        rho_even = 1.0 - (mean_var - acov[:, t + 1].mean()) / var_plus
        rho_odd = 1.0 - (mean_var - acov[:, t + 2].mean()) / var_plus
        if rho_even + rho_odd >= 0:
            rho[t + 1] = rho_even
            rho[t + 2] = rho_odd
        t += 2
    max_t = t - 2
    if rho_even > 0:
        rho[max_t + 1] = rho_even
    t = 1
    while t <= max_t - 2:
        if rho[t + 1] + rho[t + 2] > rho[t - 1] + rho[t]:
            rho[t + 1] = (rho[t - 1] + rho[t]) / 2.0
            rho[t + 2] = rho[t + 1]
        t += 2
    ess = m * n
    tau = -1.0 + 2.0 * rho[:max_t + 1].sum() + rho[max_t + 1:max_t + 2].sum()
    tau = max(tau, 1.0 / np.log10(ess))
    return float(ess / tau)


def ess_bulk(a):
    return _ess_raw(_z_scale(_split(np.asarray(a, dtype=np.float64))))


def ess_mean(a):
    return _ess_raw(_split(np.asarray(a, dtype=np.float64)))


def ess_tail(a):
    """min of the ESS of the 5 % and 95 % quantile indicators (ArviZ `_ess_tail`)."""
    a = np.asarray(a, dtype=np.float64)
    out = []
    for q in (0.05, 0.95):
        ind = (a <= np.quantile(a, q)).astype(np.float64)
        out.append(_ess_raw(_split(ind)))
    return float(min(out))


def ess_sd(a):
    """ArviZ 0.11 `_ess_sd`: min of the ESS of x and of x^2 on split chains."""
    s = _split(np.asarray(a, dtype=np.float64))
    return float(min(_ess_raw(s), _ess_raw(s ** 2)))


def mcse_sd(a):
    """ArviZ 0.11 `_mcse_sd`: sd * sqrt(e (1 - 1/ess)^(ess-1) - 1) with ess = ess_sd."""
    a = np.asarray(a, dtype=np.float64)
    e = ess_sd(a)
    if not (e == e) or e <= 1.0:
        return float("nan")
    return float(a.std(ddof=1) * np.sqrt(np.exp(1.0) * (1.0 - 1.0 / e) ** (e - 1.0) - 1.0))


def hdi(x, prob=0.94):
    x = np.sort(np.asarray(x, dtype=np.float64).ravel())
    n = len(x)
    k = int(np.floor(prob * n))
    w = x[k:] - x[:n - k]
    i = int(np.argmin(w))
    return float(x[i]), float(x[i + k])


SUMMARY_COLUMNS = ("mean", "sd", "hdi_3%", "hdi_97%", "mcse_mean", "mcse_sd", "ess_bulk", "ess_tail", "r_hat", "median")


def summary_csv(path, posterior):
    """Write `summary(posterior)` as a CSV laid out like ng_interp/ng_optPLM1.csv (first column = variable name)."""
    rows = summary(posterior)
    with open(path, "w") as f:
        f.write("," + ",".join(SUMMARY_COLUMNS) + "\n")
        for k, r in rows.items():
            f.write(k + "," + ",".join(repr(float(r[c])) for c in SUMMARY_COLUMNS) + "\n")


def summary(posterior):
    """posterior: dict var -> array [chain, draw].  Returns dict var -> row with the columns of the reference's summary
    tables (ng_interp/ng_optPLM*.csv: az.summary + median), in the same order (SUMMARY_COLUMNS)."""
    out = {}
    for k, a in posterior.items():
        a = np.asarray(a, dtype=np.float64)
        sd = a.std(ddof=1)
        em = ess_mean(a)
        lo, hi = hdi(a)
        out[k] = {"mean": float(a.mean()), "sd": float(sd), "hdi_3%": lo, "hdi_97%": hi,
                  "mcse_mean": float(sd / np.sqrt(em)) if em == em and em > 0 else float("nan"), "mcse_sd": mcse_sd(a),
                  "ess_bulk": ess_bulk(a), "ess_tail": ess_tail(a), "r_hat": rhat(a), "median": float(np.median(a))}
    return out


# ---------------------------------------------------------------------------------------- moment-based (many chains)
def moments_summary(n, mean, m2):
    """Per-chain Welford moments (n draws each; mean, m2 of shape [chains, ndim]) -> pooled mean/sd, the classic
    R-hat sqrt(((n-1)/n W + B/n)/W) and the many-chain ESS estimate M n var+/B (BDA3 eq. 11.4 with the between-chain
    variance as the variance of the chain means).  Valid when the number of chains is large."""
    mean = np.asarray(mean, dtype=np.float64)
    m2 = np.asarray(m2, dtype=np.float64)
    M = mean.shape[0]
    W = (m2 / (n - 1.0)).mean(axis=0)
    Bn = mean.var(axis=0, ddof=1)                    # B / n
    var_plus = (n - 1.0) / n * W + Bn
    return {"mean": mean.mean(axis=0), "sd": np.sqrt(var_plus), "r_hat": np.sqrt(var_plus / W),
            "ess": M * var_plus / Bn, "mcse_mean": np.sqrt(Bn / M), "chains": M, "draws_per_chain": n}


def to_inference_data(posterior, sample_stats=None, observed_data=None, attrs=None):
    """`arviz.InferenceData` of a GPU run (needs arviz; not installed in this image): the object the reference obtains from
    `az.from_pymc3(trace)` (run_age_mcmc_utils.py:419-423) -- `az.to_netcdf`, `az.summary` and the plotting scripts work
    on it directly."""
    import arviz as az
    idata = az.from_dict(posterior={k: np.asarray(v) for k, v in posterior.items()},
                         sample_stats={k: np.asarray(v) for k, v in (sample_stats or {}).items() if np.ndim(v) >= 2} or None,
                         observed_data=observed_data or None)
    for k, v in (attrs or {}).items():
        idata.posterior.attrs[k] = v
    return idata


def nested_rhat(n, mean, m2, n_super):
    """Nested R-hat of Margossian et al. (2022, "Nested R-hat: assessing the convergence of Markov chain Monte Carlo when
    running many short chains") from per-chain Welford moments: the chains [chains, ndim] are split into `n_super`
    superchains of consecutive chains; B = variance of the superchain means, W = mean over superchains of (variance of
    the chain means inside the superchain + mean within-chain variance); nR-hat = sqrt(1 + B / W).  It measures what
    matters when a POPULATION of chains is pooled (DE-MC-Z with a shared archive): whether sub-populations agree,
    not whether every single chain has mixed."""
    mean = np.asarray(mean, dtype=np.float64)
    m2 = np.asarray(m2, dtype=np.float64)
    M = mean.shape[0] // n_super
    if M < 2:
        raise ValueError("need at least two chains per superchain")
    mean = mean[:M * n_super].reshape(n_super, M, -1)
    wvar = (m2[:M * n_super] / (float(n) - 1.0)).reshape(n_super, M, -1)
    smean = mean.mean(axis=1)
    B = smean.var(axis=0, ddof=1)
    W = (mean.var(axis=1, ddof=1) + wvar.mean(axis=1)).mean(axis=0)
    return np.sqrt(1.0 + B / W)


def save_trace(path, posterior, sample_stats=None, attrs=None, observed_data=None):
    """Write a trace with the groups / variable names the reference stores through `az.to_netcdf`
    (run_age_mcmc_utils.py:425, noble_gas_mcmc.py:288): `posterior/<var>` arrays shaped [chain, draw], optional
    `sample_stats/<name>` and scalar `attrs/<name>` (e.g. sampling_time).
    A path ending in .netcdf / .nc is written as a NetCDF-4 (HDF5) file in ArviZ's InferenceData layout
    (netcdf4_writer.write_trace) -- what `az.from_netcdf` of the reference's plotting scripts opens
    (age_modeling_mcmc.post_plots.py:119-148, ng_interp/noble_gas_mcmc.compplots.py:222-233); anything else as `.npz`."""
    if str(path).endswith((".netcdf", ".nc", ".nc4")):
        from . import netcdf4_writer
        ss = {k: v for k, v in (sample_stats or {}).items() if np.ndim(v) >= 2}
        netcdf4_writer.write_trace(path, posterior, ss or None, observed_data, attrs)
        return
    out = {}
    for k, v in posterior.items():
        a = np.asarray(v, dtype=np.float64)
        if a.ndim != 2:
            raise ValueError(f"posterior[{k!r}] must be [chain, draw], got shape {a.shape}")
        out["posterior/" + k] = a
    for k, v in (sample_stats or {}).items():
        out["sample_stats/" + k] = np.asarray(v)
    for k, v in (attrs or {}).items():
        out["attrs/" + k] = np.asarray(v)
    np.savez_compressed(path, **out)


def load_trace(path):
    """Inverse of `save_trace`: returns {"posterior": {...}, "sample_stats": {...}, "attrs": {...}}.
    A NetCDF-4 / HDF5 file -- the format the reference itself writes with `az.to_netcdf` (run_age_mcmc_utils.py:425) and reads
    with `az.from_netcdf` (:434) -- is recognised by its signature and read with `netcdf4_reader.read_trace`: same layout, plus
    `observed_data`; coordinate variables (`chain`, `draw`, `*_dim_0`) are dropped and `attrs` are those of the posterior group."""
    with open(path, "rb") as fh:
        magic = fh.read(8)
    if magic == b"\x89HDF\r\n\x1a\n":
        from . import netcdf4_reader
        raw = netcdf4_reader.read_trace(path)
        tr = {"posterior": {}, "sample_stats": {}, "observed_data": {}, "attrs": {}}
        for grp in ("posterior", "sample_stats", "observed_data"):
            for k, v in raw.get(grp, {}).items():
                if k in ("chain", "draw") or k.endswith("_dim_0"):
                    continue
                tr[grp][k] = v
        for k, v in raw["attrs"].get("posterior", {}).items():
            tr["attrs"][k] = v.reshape(-1)[0].item() if isinstance(v, np.ndarray) and v.size == 1 else v
        return tr
    tr = {"posterior": {}, "sample_stats": {}, "attrs": {}}
    with np.load(path) as z:
        for k in z.files:
            grp, _, name = k.partition("/")
            if grp not in tr or not name:
                raise ValueError(f"{path}: unexpected entry {k!r}")
            a = z[k]
            tr[grp][name] = a.item() if grp == "attrs" and a.ndim == 0 else a
    return tr
