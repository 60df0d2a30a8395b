"""TEST INFRASTRUCTURE ONLY -- numpy restatement of the sampler semantics (pymc3 3.11.2 DEMetropolisZ + metrop_select +
transforms) driven by the same counter-based Philox4x32-10 stream as the device sampler, so that device
trajectories can be checked step by step.  pymc3 is a third-party dependency absent from /root/reference: the
*algorithm* is restated from its published source; parity with the reference is statistical (posterior summaries in
ng_interp/ng_optPLM*.csv), see DESIGN.md.  Never imported by the product package."""
import numpy as np
from scipy.special import betaln, gammaln

M0, M1, W0, W1 = 0xD2511F53, 0xCD9E8D57, 0x9E3779B9, 0xBB67AE85
MASK = 0xFFFFFFFF


def philox4x32_10(ctr, key):
    """Salmon et al. SC'11; ctr 4 x uint32, key 2 x uint32 (python ints)."""
    c = [int(x) & MASK for x in ctr]
    k = [int(x) & MASK for x in key]
    for _ in range(10):
        p0 = M0 * c[0]
        p1 = M1 * c[2]
        c = [((p1 >> 32) ^ c[1] ^ k[0]) & MASK, p1 & MASK, ((p0 >> 32) ^ c[3] ^ k[1]) & MASK, p0 & MASK]
        k = [(k[0] + W0) & MASK, (k[1] + W1) & MASK]
    return c


def chain_rng(seed, chain, step, purpose):
    return philox4x32_10([chain & MASK, (chain >> 32) & MASK, step & MASK, purpose | (((step >> 32) & 0xFFFF) << 16)],
                         [seed & MASK, (seed >> 32) & MASK])


def u01(a, b):
    x = (a << 32) | b
    return ((x >> 11) + 0.5) / 9007199254740992.0


def softplus(y):
    return y + np.log1p(np.exp(-y)) if y > 0 else np.log1p(np.exp(y))


def transform_dim(pr, x):
    """-> (natural value, log prior + log |Jacobian|)."""
    k = pr["kind"]
    if k == "uniform":
        sp = softplus(-x)
        return pr["p0"] + (pr["p1"] - pr["p0"]) * np.exp(-sp), -2.0 * sp - x
    if k == "beta":
        sp = softplus(-x)
        return (pr["lo"] + (pr["hi"] - pr["lo"]) * np.exp(-sp),
                pr["p0"] * (-sp) + pr["p1"] * (-(x + sp)) - betaln(pr["p0"], pr["p1"]))
    if k == "normal":
        z = (x - pr["p0"]) / pr["p1"]
        return x, -0.5 * np.log(2 * np.pi * pr["p1"] ** 2) - 0.5 * z * z
    v = np.exp(x)
    return v, 0.5 * np.log(2 / np.pi) - np.log(pr["p0"]) - v * v / (2 * pr["p0"] ** 2) + x


def test_point(priors):
    q = []
    for pr in priors:
        if pr["kind"] == "uniform":
            q.append(0.0)
        elif pr["kind"] == "beta":
            m = pr["p0"] / (pr["p0"] + pr["p1"])
            q.append(np.log(m / (1 - m)))
        elif pr["kind"] == "normal":
            q.append(pr["p0"])
        else:
            q.append(np.log(pr["p0"] * np.sqrt(2 / np.pi)))
    return np.array(q)


def tune_factor(r):
    if r < 0.001:
        return 0.1
    if r < 0.05:
        return 0.5
    if r < 0.2:
        return 0.9
    if r > 0.95:
        return 10.0
    if r > 0.75:
        return 2.0
    if r > 0.5:
        return 1.1
    return 1.0


def studentt_logp(obs, mu, sd, nu):
    lam = sd ** -2.0
    return float(np.sum(gammaln((nu + 1) / 2) - gammaln(nu / 2) + 0.5 * np.log(lam / (nu * np.pi))
                        - (nu + 1) / 2 * np.log1p(lam * (obs - mu) ** 2 / nu)))


def normal_logp(obs, mu, sd):
    return float(np.sum(-0.5 * np.log(2 * np.pi * sd ** 2) - (obs - mu) ** 2 / (2 * sd ** 2)))


def run_chain(priors, logp_model, nsteps, seed, chain, tune_steps=0, tune_interval=1000, scaling=0.001, lamb=None,
              tune_drop_fraction=0.9, de_mcz=True, q0=None):
    """One chain of DE-MC-Z.  logp_model(values: dict target -> natural value) -> log likelihood.
    Returns (q trajectory [nsteps, nd], logp [nsteps], accepted [nsteps])."""
    nd = len(priors)
    lamb = 2.38 / np.sqrt(2 * nd) if lamb is None else lamb

    def full_logp(q):
        vals, lp = {}, 0.0
        for d, pr in enumerate(priors):
            v, l = transform_dim(pr, q[d])
            vals[pr["target"]] = v
            lp += l
        with np.errstate(all="ignore"):
            return lp + logp_model(vals)

    q = test_point(priors) if q0 is None else np.array(q0, dtype=np.float64)
    logp = full_logp(q)
    hist, hist_start = [], 0
    acc_win = 0
    Q, LP, AC = [], [], []
    for i in range(nsteps):
        tuning = i < tune_steps
        if i == tune_steps and tune_steps > 0:                      # stop_tuning
            hist_start += int(tune_drop_fraction * (i - hist_start))
        if tuning and i > 0 and i % tune_interval == 0:
            lamb *= tune_factor(acc_win / float(tune_interval))
            acc_win = 0
        nvalid = i - hist_start
        sel = chain_rng(seed, chain, i, 0x100)
        qn = q.copy()
        if de_mcz and nvalid > 1:
            iz1 = (sel[0] * nvalid) >> 32
            iz2 = (sel[1] * (nvalid - 1)) >> 32
            if iz2 >= iz1:
                iz2 += 1
            qn = qn + lamb * (hist[i - nvalid + iz1] - hist[i - nvalid + iz2])
        for d in range(nd):
            e = chain_rng(seed, chain, i, d)
            qn[d] += (2.0 * u01(e[0], e[1]) - 1.0) * scaling
        lpn = full_logp(qn)
        delta = lpn - logp
        acc = bool(np.isfinite(delta) and np.log(u01(sel[2], sel[3])) < delta)
        if acc:
            q, logp = qn, lpn
            acc_win += 1
        hist.append(q.copy())
        Q.append(q.copy()); LP.append(logp); AC.append(acc)
    return np.array(Q), np.array(LP), np.array(AC)
