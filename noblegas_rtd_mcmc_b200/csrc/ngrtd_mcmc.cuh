// ngrtd_mcmc.cuh -- device-resident batched Metropolis sampler (K5) and per-chain running statistics (K6).
//
// Restates the sampler the reference configures through the third-party pymc3 3.11.2 (absent from /root/reference;
// call sites age_ens_runs_mcmc/run_age_mcmc_utils.py:286-344,388-396,412-417 and ng_interp/noble_gas_mcmc.py:224-266,
// 408-415; semantics in SURVEY.md App. B):
//   * free variables live in pymc3's transformed space (interval / log-odds / log transforms, log-Jacobian included)
//   * DEMetropolisZ: q = q0 + lamb*(z1 - z2) + U(-1,1)^nd * scaling with z1 != z2 drawn from the chain's OWN history
//     (every step, rejected or not, appends the current state), lamb = 2.38/sqrt(2 nd) tuned by acceptance rate every
//     tune_interval steps, 90 % of the history dropped when tuning stops; metrop_select accepts iff
//     isfinite(delta) and log U < delta.
//   * random numbers: Philox4x32-10, key = seed, counter = (global chain id, step, purpose); a chain's trajectory is a
//     pure function of (seed, global chain id), so results do not depend on how chains are sharded over GPUs.
// Whole steps (propose -> transform/prior -> forward -> likelihood -> accept -> history/trace/statistics) run inside
// one persistent kernel; a launch advances every chain by `nsteps` steps with no host round trip.
#pragma once
#include "ngrtd_ce.cuh"
#include "ngrtd_forward.cuh"

namespace ngrtd {

#ifndef NGRTD_ND_MAX
#define NGRTD_ND_MAX 10
#endif
constexpr int ND_MAX = NGRTD_ND_MAX;      // sampler dimensions
constexpr int NVAL = 12;        // value registers: 0..10 = ForwardMod.p_dict slots, 11 = nu_ (raw, in [0,1])
constexpr int VAL_NU = 11;
// per-chain shared-memory record of the age kernel (doubles): qs[ndr] | qp[ndr] | vals[NVAL] | 8 scalars, ndr = ND_SMALL for
// samplers of up to 8 dimensions, else ND_MAX (r2 session 3: with ND_MAX the 256 records of a CTA took 80 KB and left no room
// for the 2,048-entry exp table next to 840 resident lags; 8 dimensions -> 72 KB)
constexpr int ND_SMALL = 8;
__host__ __device__ constexpr int ch_rec_doubles(int ndr) { return 2 * ndr + NVAL + 8; }
enum PriorKind : int { PR_UNIFORM = 0, PR_BETA = 1, PR_NORMAL = 2, PR_HALFNORMAL = 3 };

struct PriorDev {
    int kind;
    int target;          // value register written by this dimension
    double p0, p1;       // uniform: a, b | beta: alpha, beta | normal: mu, sigma | halfnormal: sigma, -
    double lo, hi;       // beta: affine map of the [0,1] variable onto [lo, hi]
    double c;            // normalising constant of the log density
};

struct SamplerView {
    int nd;
    PriorDev pr[ND_MAX];
    int model;               // 0: age model (convolution plan), 1: noble-gas closed-equilibrium model
    int lik_kind;            // 0 normal, 1 student-t
    int nu_sampled;          // 1: nu = nu_lo + (nu_hi - nu_lo) * val[VAL_NU]
    double nu_lo, nu_hi, nu_fixed;
    int ntr;
    double obs[MAX_TRACER], isd[MAX_TRACER], lc[MAX_TRACER];   // observations, 1/sd, per-tracer likelihood constants
    int f2_from_f1;          // f2 = 1 - f1 (run_age_mcmc_utils.py:304)
    unsigned int sampled_mask;   // bit t set: value register t is driven by a sampler dimension
    int proposal_dist;       // 0 uniform(-1,1), 1 normal(0,1)
    int de_mcz;              // 1: DE-MC-Z proposals, 0: plain random walk
    int tune_target;         // 0 lambda, 1 scaling
    int tune_interval;
    unsigned long long seed;
    long long chain_offset;  // global id of local chain 0 (shard offset)
    long long B;
    double* q;               // [B, nd] transformed state
    double* logp;            // [B]
    double* lamb;            // [B]
    double* scal;            // [B]
    int* acc_win;            // [B] accepted since the last tuning point
    long long* acc_tot;      // [B] accepted since creation
    double* hist;            // [cap, B, nd] ring of past states
    int hist_cap;
    double* wf_mean;         // [B, nd] running mean of the natural values (Welford)
    double* wf_m2;           // [B, nd] running sum of squared deviations
    // per-group observations (config 4: wells x ensemble members): tables [G, ntr]; group = global chain id / cpg
    const double* g_obs;
    const double* g_isd;
    const double* g_lc;
    long long cpg;           // chains per group (0: one observation vector for all chains, obs/isd/lc above)
    long long pool;          // > 0: DE-MC-Z with a SHARED archive -- z1, z2 come from the histories of the `pool` consecutive
                             // global chains of this chain's population (ter Braak & Vrugt 2008), see de_select()
    GasList gases;           // noble-gas model: modelled gases
    double val_defaults[NVAL];   // value registers not driven by a sampler dimension (p_dict defaults)
};

struct RunArgs {
    long long step0;         // global index of the first step of this launch
    int nsteps;
    int mode;                // 0: Metropolis steps, 1: evaluate logp of the current state (initialisation)
    int tune;                // 1: tuning phase
    long long hist_start;    // logical index of the oldest valid history entry
    double* trace;           // [ndraw, B, nd] natural values (or nullptr)
    int thin;
    long long draw0;         // draws recorded before this launch (trace row / Welford count offset)
    int record;              // 1: record trace rows and statistics
};

// ---------------------------------------------------------------- Philox4x32-10 (Salmon et al., SC'11)
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
    const unsigned int M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
    for (int i = 0; i < 10; i++) {
        unsigned int hi0 = __umulhi(M0, c.x), lo0 = M0 * c.x;
        unsigned int hi1 = __umulhi(M1, c.z), lo1 = M1 * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += W0;
        k.y += W1;
    }
    return c;
}
__device__ __forceinline__ uint4 chain_rng(unsigned long long seed, long long chain, long long step, unsigned int purpose) {
    uint4 c = make_uint4((unsigned int)chain, (unsigned int)((unsigned long long)chain >> 32), (unsigned int)step,
                         purpose | ((unsigned int)((unsigned long long)step >> 32) << 16));
    uint2 k = make_uint2((unsigned int)seed, (unsigned int)(seed >> 32));
    return philox4x32_10(c, k);
}
__device__ __forceinline__ double u01(unsigned int a, unsigned int b) {   // uniform on (0,1), 53 bits, never 0 or 1
    unsigned long long x = ((unsigned long long)a << 32) | b;
    return ((double)(x >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
constexpr unsigned int RNG_SELECT = 0x100u;    // purpose of the (iz1, iz2, accept-uniform) draw
constexpr unsigned int RNG_POOL = 0x101u;      // purpose of the partner-chain draw of the shared-archive mode

// Which two history entries form the DE-MC-Z difference of `chain` at step i.
// Default (pymc3 3.11.2 DEMetropolisZ): two distinct entries of the chain's OWN history, uniformly among the nvalid newest.
// Shared archive (sv.pool > 0; the original DE-MC-Z of ter Braak & Vrugt 2008, where Z is the archive of ALL chains of a
// population): each of z1, z2 is an entry of a uniformly drawn chain of the population.  A chain that only ever sees its
// own past cannot propose a jump to a mode it has never visited; with the shared archive the difference vectors span the
// modes the population has found (BASELINE config 4: 256 chains per (well, ensemble member)).  To stay deterministic the
// (Every ~10th proposal of this mode takes gamma = 1 instead of the tuned lambda -- the mode-jumping move of DE-MC.)
// partners' entries are taken from what was complete when this LAUNCH started and cannot be overwritten during it:
// logical entries [max(hist_start, step0 + nsteps - cap), step0); before such a window exists the own history is used.
// Populations must not straddle shards (partners are clipped to the local chains).
// history ring: logical entry e lives in slot e % cap
__device__ __forceinline__ size_t hist_off(const SamplerView& sv, long long e, long long chain) {
    return ((size_t)(e % sv.hist_cap) * (size_t)sv.B + (size_t)chain) * (size_t)sv.nd;
}

struct DeSel { size_t o1, o2; bool use, jump; };
__device__ __forceinline__ DeSel de_select(const SamplerView& sv, const RunArgs& ra, long long chain, long long gchain,
                                           long long i, uint4 sel) {
    DeSel r{0, 0, false, false};
    long long nvalid = i - ra.hist_start;
    if (nvalid > sv.hist_cap) nvalid = sv.hist_cap;
    if (!(ra.mode == 0 && sv.de_mcz && nvalid > 1)) return r;
    r.use = true;
    if (sv.pool > 0) {
        long long lo = ra.step0 + ra.nsteps - sv.hist_cap;
        if (lo < ra.hist_start) lo = ra.hist_start;
        const long long nw = ra.step0 - lo;
        if (nw > 1) {
            long long g0 = (gchain / sv.pool) * sv.pool - sv.chain_offset;
            long long g1 = g0 + sv.pool;
            if (g0 < 0) g0 = 0;
            if (g1 > sv.B) g1 = sv.B;
            const unsigned int np = (unsigned int)(g1 - g0);
            const uint4 s2 = chain_rng(sv.seed, gchain, i, RNG_POOL);
            const long long c1 = g0 + __umulhi(s2.x, np), c2 = g0 + __umulhi(s2.y, np);
            const long long e1 = lo + __umulhi(sel.x, (unsigned int)nw), e2 = lo + __umulhi(sel.y, (unsigned int)nw);
            r.o1 = hist_off(sv, e1, c1);
            r.o2 = hist_off(sv, e2, c2);
            // every ~10th proposal uses gamma = 1 instead of the tuned lambda: the full difference of two archive members is
            // what carries a chain from one mode to another (ter Braak 2006, section 2)
            r.jump = s2.z < 429496730u;
            return r;
        }
    }
    unsigned int iz1 = __umulhi(sel.x, (unsigned int)nvalid);
    unsigned int iz2 = __umulhi(sel.y, (unsigned int)(nvalid - 1));
    if (iz2 >= iz1) iz2++;
    r.o1 = hist_off(sv, i - nvalid + iz1, chain);
    r.o2 = hist_off(sv, i - nvalid + iz2, chain);
    return r;
}

// ---------------------------------------------------------------- priors and transforms (pymc3 3.11.2)
// transformed coordinate x -> natural value v; returns log prior density + log |Jacobian|.
// sigmoid and its log share one exp and one log1p:  e = exp(-|x|), log sig(x) = min(x,0) - log1p(e), sig(x) = [x>=0 ? 1 : e]/(1+e)
__device__ __forceinline__ double transform_dim(const PriorDev& pr, double x, double& v) {
    switch (pr.kind) {
        case PR_UNIFORM:        // interval transform: v = a + (b-a) sigmoid(x); logp + jac = log sig(x) + log sig(-x)
        case PR_BETA: {         // log-odds transform; natural value mapped affinely onto [lo, hi]
            double e = exp(-fabs(x));
            double l1 = log1p(e);
            double sig = ((x >= 0.0) ? 1.0 : e) / (1.0 + e);
            double lsp = fmin(x, 0.0) - l1;          // log sigmoid(x)
            double lsm = lsp - x;                    // log sigmoid(-x)
            if (pr.kind == PR_UNIFORM) {
                v = pr.p0 + (pr.p1 - pr.p0) * sig;
                return lsp + lsm;
            }
            v = pr.lo + (pr.hi - pr.lo) * sig;
            return pr.p0 * lsp + pr.p1 * lsm - pr.c;
        }
        case PR_NORMAL: {
            double z = (x - pr.p0) / pr.p1;
            v = x;
            return pr.c - 0.5 * z * z;
        }
        default: {              // half-normal, log transform
            v = exp(x);
            return pr.c - v * v / (2.0 * pr.p0 * pr.p0) + x;
        }
    }
}

// pymc3 step_methods/metropolis.py tune(): rescale by acceptance rate
__device__ __forceinline__ double tune_factor(double acc_rate) {
    if (acc_rate < 0.001) return 0.1;
    if (acc_rate < 0.05) return 0.5;
    if (acc_rate < 0.2) return 0.9;
    if (acc_rate > 0.95) return 10.0;
    if (acc_rate > 0.75) return 2.0;
    if (acc_rate > 0.5) return 1.1;
    return 1.0;
}

__device__ __forceinline__ double proposal_noise(int dist, uint4 e) {
    double u = u01(e.x, e.y);
    if (dist == 0) return 2.0 * u - 1.0;
    double u2 = u01(e.z, e.w);
    return sqrt(-2.0 * log(u)) * cospi(2.0 * u2);
}

// Welford update + trace row for one dimension
__device__ __forceinline__ void record_dim(const SamplerView& sv, const RunArgs& ra, long long chain, int d, double v,
                                           long long draw) {
    if (ra.trace) ra.trace[((size_t)(draw - ra.draw0) * (size_t)sv.B + (size_t)chain) * sv.nd + d] = v;
    size_t o = (size_t)chain * sv.nd + d;
    double n = (double)(draw + 1);
    double mean = sv.wf_mean[o];
    double dl = v - mean;
    mean += dl / n;
    sv.wf_mean[o] = mean;
    sv.wf_m2[o] += dl * (v - mean);
}

// ---------------------------------------------------------------- age model: one warp = NT tiles of 8 chains
// TAIL: the constant-tail code (closed form / quadrature) is compiled in; the launcher picks the instantiation by plan
// TB: exp table of the dispersion weights (ExpCfg<TB>): 11 when the launcher finds room for it next to the resident lag tables
//     (one DFMA less per dispersion weight: 0.1324 -> 0.1256 ms per step on the cfg-3 sampler), else 7
// NDR: dimensions a per-chain record has room for (compile time: a run-time record stride measured 4 % slower -- it stays
//      live across the lag loop of a kernel that is at its register limit)
template <int C1, int C2, bool DYN, int NT, int UA, int MAXW, bool TAIL, int TB, int NDR>
__global__ void __launch_bounds__(MAXW * 32, 1)
k_mcmc_age(PlanView pv, SamplerView sv, RunArgs ra, int lc_cap) {
    FwdCta<C1, C2, DYN, NT, UA, TB, TAIL ? 1 : 0> cta(pv);
    constexpr int ndr = NDR, CH_REC = ch_rec_doubles(NDR);
    const int rec_base = cta.setup(lc_cap);
    const int lane = cta.lane, j = lane & 3, r = lane >> 2;
    // prior table in shared memory (dynamic indexing of kernel parameters would be demoted to local memory)
    const int pr_off = rec_base + blockDim.x / 32 * NT * 8 * CH_REC;
    PriorDev* spr = reinterpret_cast<PriorDev*>(ngrtd_smem + pr_off);
    if (threadIdx.x < sv.nd) spr[threadIdx.x] = sv.pr[threadIdx.x];
    __syncthreads();
    const int rec_warp = rec_base + cta.warp * NT * 8 * CH_REC;
    const long long B = sv.B;
    const long long nunits = (B + NT * 8 - 1) / (NT * 8);
    double ob0[2], is0[2], lc0[2];              // shared observation vector: the lane's two tracers (j, j+4)
#pragma unroll
    for (int q = 0; q < 2; q++) {
        int tr = min(j + 4 * q, MAX_TRACER - 1);
        ob0[q] = sv.obs[tr]; is0[q] = sv.isd[tr]; lc0[q] = sv.lc[tr];
    }

    for (cta.sched_begin(nunits); cta.sched_valid(); cta.sched_next()) {
        const long long u = cta.unit;
        const bool active = cta.active, lockstep = cta.lockstep;
        long long chain[NT];
        int rec[NT];
        bool ok[NT];
        // ---- load the chain state into shared memory: qs | qp | vals | scalars(logp, lamb, scal, lpsum, uacc, acc_win, acc_tot)
#pragma unroll
        for (int t = 0; t < NT; t++) {
            chain[t] = (u * NT + t) * 8 + r;
            ok[t] = active && chain[t] < B;
            long long cl = (active && chain[t] < B) ? chain[t] : 0;
            rec[t] = rec_warp + (t * 8 + r) * CH_REC;
            for (int d = j; d < sv.nd; d += 4) ngrtd_smem[rec[t] + d] = sv.q[cl * sv.nd + d];
#pragma unroll
            for (int v = 0; v < NVAL / 4; v++) ngrtd_smem[rec[t] + 2 * ndr + 4 * v + j] = sv.val_defaults[4 * v + j];
            if (j == 0) {
                double* sc = ngrtd_smem + rec[t] + 2 * ndr + NVAL;
                sc[0] = sv.logp[cl];
                sc[1] = sv.lamb[cl];
                sc[2] = sv.scal[cl];
                sc[5] = (double)sv.acc_win[cl];
                sc[6] = (double)sv.acc_tot[cl];
            }
        }
        __syncwarp();
        const int nsteps = ra.mode == 1 ? 1 : ra.nsteps;
        for (int s = 0; s < nsteps; s++) {
            const long long i = ra.step0 + s;
            ChainPar par[NT];
            double nu[NT];
#pragma unroll
            for (int t = 0; t < NT; t++) {
                double* qs = ngrtd_smem + rec[t];
                double* qp = qs + ndr;
                double* vals = qs + 2 * ndr;
                double* sc = vals + NVAL;
                const long long gchain = sv.chain_offset + chain[t];
                // -- tuning point (pymc3 DEMetropolisZ.astep: rescale before proposing)
                if (ra.mode == 0 && ra.tune && i > 0 && (i % sv.tune_interval) == 0 && j == 0) {
                    double f = tune_factor(sc[5] / (double)sv.tune_interval);
                    if (sv.tune_target == 0) sc[1] *= f; else sc[2] *= f;
                    sc[5] = 0.0;
                }
                __syncwarp();
                // -- proposal
                uint4 sel = chain_rng(sv.seed, gchain, i, RNG_SELECT);
                const DeSel de = de_select(sv, ra, ok[t] ? chain[t] : 0, gchain, i, sel);
                const bool use_de = de.use;
                const double lamb = de.jump ? 1.0 : sc[1], scal = sc[2];
                double lps = 0.0;
                for (int d = j; d < sv.nd; d += 4) {
                    double qn = qs[d];
                    if (ra.mode == 0) {
                        double eps = proposal_noise(sv.proposal_dist, chain_rng(sv.seed, gchain, i, (unsigned int)d));
                        if (use_de && ok[t]) qn += lamb * (sv.hist[de.o1 + d] - sv.hist[de.o2 + d]);
                        qn += eps * scal;
                    }
                    qp[d] = qn;
                    double v;
                    lps += transform_dim(spr[d], qn, v);
                    vals[spr[d].target] = v;
                }
                lps += __shfl_xor_sync(0xffffffffu, lps, 1);
                lps += __shfl_xor_sync(0xffffffffu, lps, 2);
                if (j == 0) { sc[3] = lps; sc[4] = u01(sel.z, sel.w); }
                __syncwarp();
                // -- natural parameters of the forward model
                ChainPar& p = par[t];
                p.tau1 = vals[0];
                p.tau2 = vals[1];
                p.f1 = vals[2];
                p.f2 = sv.f2_from_f1 ? 1.0 - vals[2] : vals[3];
                p.eta1 = pv.eta1_is_one ? 1.0 : vals[4];
                p.eta2 = pv.eta2_is_one ? 1.0 : vals[5];
                p.D1 = vals[6];
                p.D2 = vals[7];
                p.log10J = vals[8];
                p.lam_cfc = (sv.sampled_mask & (1u << 9)) ? LN2 / vals[9] : 0.0;
                p.lamsf6 = vals[10];
                nu[t] = sv.nu_sampled ? sv.nu_lo + (sv.nu_hi - sv.nu_lo) * vals[VAL_NU] : sv.nu_fixed;
            }
            double val[NT][2];
            cta.eval(par, active, lockstep, val);
            if (!active) continue;
#pragma unroll
            for (int t = 0; t < NT; t++) {
                double* qs = ngrtd_smem + rec[t];
                double* qp = qs + ndr;
                double* sc = qs + 2 * ndr + NVAL;
                double ob[2] = {ob0[0], ob0[1]}, is[2] = {is0[0], is0[1]}, lc[2] = {lc0[0], lc0[1]};
                if (sv.cpg > 0) {                    // config 4: this chain's own observation row
                    const long long grp = (sv.chain_offset + (ok[t] ? chain[t] : 0)) / sv.cpg;
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        long long o = grp * pv.ntracer + min(j + 4 * q, pv.ntracer - 1);
                        ob[q] = sv.g_obs[o]; is[q] = sv.g_isd[o]; lc[q] = sv.g_lc[o];
                    }
                }
                double ll = lik_reduce(sv.lik_kind, pv.ntracer, j, val[t], ob, is, lc, nu[t]);
                double lpn = sc[3] + ll;
                double delta = lpn - sc[0];
                bool acc = ra.mode == 1 || (isfinite(delta) && log(sc[4]) < delta);    // metrop_select
                __syncwarp();
                if (acc) {
                    for (int d = j; d < sv.nd; d += 4) qs[d] = qp[d];
                    if (j == 0) { sc[0] = lpn; if (ra.mode == 0) { sc[5] += 1.0; sc[6] += 1.0; } }
                }
                __syncwarp();
                if (ra.mode == 0 && ok[t]) {
                    size_t ho = hist_off(sv, i, chain[t]);                              // history.append(q_new)
                    for (int d = j; d < sv.nd; d += 4) sv.hist[ho + d] = qs[d];
                    if (ra.record && ((i - ra.step0) % ra.thin) == 0) {
                        long long draw = ra.draw0 + (i - ra.step0) / ra.thin;
                        for (int d = j; d < sv.nd; d += 4) {
                            double v;
                            transform_dim(spr[d], qs[d], v);
                            record_dim(sv, ra, chain[t], d, v, draw);
                        }
                    }
                }
            }
        }
        // ---- store the chain state
        __syncwarp();
#pragma unroll
        for (int t = 0; t < NT; t++) {
            if (!ok[t]) continue;
            for (int d = j; d < sv.nd; d += 4) sv.q[chain[t] * sv.nd + d] = ngrtd_smem[rec[t] + d];
            if (j == 0) {
                double* sc = ngrtd_smem + rec[t] + 2 * ndr + NVAL;
                sv.logp[chain[t]] = sc[0];
                sv.lamb[chain[t]] = sc[1];
                sv.scal[chain[t]] = sc[2];
                sv.acc_win[chain[t]] = (int)sc[5];
                sv.acc_tot[chain[t]] = (long long)sc[6];
            }
        }
        __syncwarp();
    }
    if (cta.pending) cta.wait_chunk();      // warps without work must not exit under an in-flight bulk copy
}

// ---------------------------------------------------------------- noble-gas CE model: one chain per thread
// value registers (ng_interp/noble_gas_mcmc.py:224-250): 0 log10 Ae, 1 log10 F, 2 E, 3 m, 4 b, 11 nu_;  T = (E - b)/m
#ifndef NGRTD_NG_MINBLOCKS
#define NGRTD_NG_MINBLOCKS 1
#endif
#ifndef NGRTD_NO_AUX_KERNELS   // compiled into part 0 only (ngrtd_api.cu, build partitioning)
__global__ void __launch_bounds__(64, NGRTD_NG_MINBLOCKS) k_mcmc_ng(SamplerView sv, RunArgs ra) {
    const long long chain = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (chain >= sv.B) return;
    const long long gchain = sv.chain_offset + chain;
    double qs[ND_MAX], qp[ND_MAX], vals[NVAL];
    double vn[ND_MAX], vp[ND_MAX];          // natural values of the current state / of the proposal (recorded without re-transforming)
    for (int d = 0; d < sv.nd; d++) {
        qs[d] = sv.q[chain * sv.nd + d];
        transform_dim(sv.pr[d], qs[d], vn[d]);
    }
    for (int v = 0; v < NVAL; v++) vals[v] = sv.val_defaults[v];
    double logp = sv.logp[chain], lamb = sv.lamb[chain], scal = sv.scal[chain];
    int acc_win = sv.acc_win[chain];
    long long acc_tot = sv.acc_tot[chain];
    const int nsteps = ra.mode == 1 ? 1 : ra.nsteps;
    for (int s = 0; s < nsteps; s++) {
        const long long i = ra.step0 + s;
        if (ra.mode == 0 && ra.tune && i > 0 && (i % sv.tune_interval) == 0) {
            double f = tune_factor((double)acc_win / (double)sv.tune_interval);
            if (sv.tune_target == 0) lamb *= f; else scal *= f;
            acc_win = 0;
        }
        const uint4 sel = chain_rng(sv.seed, gchain, i, RNG_SELECT);
        const DeSel de = de_select(sv, ra, chain, gchain, i, sel);
        const bool use_de = de.use;
        const size_t o1 = de.o1, o2 = de.o2;
        double lps = 0.0;
        for (int d = 0; d < sv.nd; d++) {
            double qn = qs[d];
            if (ra.mode == 0) {
                double eps = proposal_noise(sv.proposal_dist, chain_rng(sv.seed, gchain, i, (unsigned int)d));
                if (use_de) qn += (de.jump ? 1.0 : lamb) * (sv.hist[o1 + d] - sv.hist[o2 + d]);
                qn += eps * scal;
            }
            qp[d] = qn;
            double v;
            lps += transform_dim(sv.pr[d], qn, v);
            vp[d] = v;
            vals[sv.pr[d].target] = v;
        }
        // forward model: ce_exc_wrapper (ng_interp/noble_gas_mcmc.py:205-213) with T from the lapse-rate line (:240)
        double Ae = exp10(vals[0]), F = exp10(vals[1]), E = vals[2];
        double T = (E - vals[4]) / vals[3];
        double P = ce_lapse_rate_step(E);
        const CeStep cs = ce_step(T, P);
        double nu = sv.nu_sampled ? sv.nu_lo + (sv.nu_hi - sv.nu_lo) * vals[VAL_NU] : sv.nu_fixed;
        double cst = sv.lik_kind == 1 ? lik_studentt_const(nu) : 0.0;
        double ll = 0.0;
        const long long grp = sv.cpg > 0 ? gchain / sv.cpg : 0;
        for (int g = 0; g < sv.gases.n; g++) {
            double mu = ce_exc_step(sv.gases.id[g], cs, E, T, Ae, F, P);
            double ob = sv.cpg > 0 ? sv.g_obs[grp * sv.gases.n + g] : sv.obs[g];
            double is = sv.cpg > 0 ? sv.g_isd[grp * sv.gases.n + g] : sv.isd[g];
            double lc = sv.cpg > 0 ? sv.g_lc[grp * sv.gases.n + g] : sv.lc[g];
            ll += sv.lik_kind == 1 ? lik_term_studentt(ob, mu, is, lc, nu, cst) : lik_term_normal(ob, mu, is, lc);
        }
        double lpn = lps + ll;
        double delta = lpn - logp;
        bool acc = ra.mode == 1 || (isfinite(delta) && log(u01(sel.z, sel.w)) < delta);
        if (acc) {
            for (int d = 0; d < sv.nd; d++) { qs[d] = qp[d]; vn[d] = vp[d]; }
            logp = lpn;
            if (ra.mode == 0) { acc_win++; acc_tot++; }
        }
        if (ra.mode == 0) {
            size_t ho = hist_off(sv, i, chain);
            for (int d = 0; d < sv.nd; d++) sv.hist[ho + d] = qs[d];
            if (ra.record && ((i - ra.step0) % ra.thin) == 0) {
                long long draw = ra.draw0 + (i - ra.step0) / ra.thin;
                for (int d = 0; d < sv.nd; d++) record_dim(sv, ra, chain, d, vn[d], draw);
            }
        }
    }
    for (int d = 0; d < sv.nd; d++) sv.q[chain * sv.nd + d] = qs[d];
    sv.logp[chain] = logp;
    sv.lamb[chain] = lamb;
    sv.scal[chain] = scal;
    sv.acc_win[chain] = acc_win;
    sv.acc_tot[chain] = acc_tot;
}
#endif

// ---------------------------------------------------------------- noble-gas CE model, register-resident (r2)
// Same arithmetic, same Philox counters and therefore the same trajectories as k_mcmc_ng, for a compile-time number of
// dimensions: every per-dimension array has constant indices after unrolling and lives in registers (k_mcmc_ng keeps a 416-byte
// local-memory frame because its loops run to the run-time sv.nd), and the history gather is software-pipelined: Philox is
// counter based, so the selection of step s + 1 is computed at the top of step s and its two rows start an ASYNCHRONOUS copy
// into the thread's shared-memory slot there (cp.async, SASS LDGSTS: no destination register, so nothing can stall on it and
// nothing is spilled); they are read at the END of the step (difference vector of the next proposal), after the whole CE model.
// r1 / r2 profiles: the subtraction z1 - z2 at the top of the step carried 12-14 % of all stall samples; cache hints
// (prefetch.global.L2, touch loads) did not move it and a register prefetch was spilled by ptxas right behind the loads
// (profiles/r2_notes.md).
// Hazard: the next step may select the entry this step appends (own-history mode); that row is then the chain's final state
// of this step, which is in registers.
// >= 7 blocks of 64 threads per SM = 448 chains: one wave for 65,536 chains on 148 SMs (ptxas settles on 128 registers; a
// 142-register build without any spill holds 6 blocks and is 1.5x slower at that batch)
#ifndef NGRTD_NG_R_MINBLOCKS
#define NGRTD_NG_R_MINBLOCKS 8
#endif
template <int ND>
__global__ void __launch_bounds__(64, NGRTD_NG_R_MINBLOCKS) k_mcmc_ng_r(SamplerView sv, RunArgs ra) {
    __shared__ double zrow[2 * ND * 64];           // [row 1 | row 2][dim][thread]: conflict-free
    __shared__ double prop[3 * ND * 64];           // the proposal (transformed | natural) between its construction and the accept
                                                   // decision, and the natural values of the current state (read only when a
                                                   // draw is recorded): 6 ND registers less across the CE model
    const long long chain = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (chain >= sv.B) return;
    const long long gchain = sv.chain_offset + chain;
    double* z1s = zrow + threadIdx.x;
    double* z2s = zrow + ND * 64 + threadIdx.x;
    const unsigned int z1a = (unsigned int)__cvta_generic_to_shared(z1s), z2a = (unsigned int)__cvta_generic_to_shared(z2s);
    double* qp = prop + threadIdx.x;               // qp[d * 64]
    double* vp = prop + ND * 64 + threadIdx.x;
    double* vn = prop + 2 * ND * 64 + threadIdx.x;
    // Welford accumulators of the chain: read once per launch, updated in shared memory, written back once (k_mcmc_ng
    // reads and writes them in global memory every recorded step: 2/3 of its DRAM traffic)
    __shared__ double wfs[2 * ND * 64];
    double* wmean = wfs + threadIdx.x;
    double* wm2 = wfs + ND * 64 + threadIdx.x;
    const bool rec = ra.mode == 0 && ra.record;
    if (rec) {
#pragma unroll
        for (int d = 0; d < ND; d++) { wmean[d * 64] = sv.wf_mean[chain * ND + d]; wm2[d * 64] = sv.wf_m2[chain * ND + d]; }
    }
    double qs[ND];
#pragma unroll
    for (int d = 0; d < ND; d++) {
        qs[d] = sv.q[chain * ND + d];
        double v;
        transform_dim(sv.pr[d], qs[d], v);
        vn[d * 64] = v;
    }
    double logp = sv.logp[chain], lamb = sv.lamb[chain], scal = sv.scal[chain];
    int acc_win = sv.acc_win[chain];
    long long acc_tot = sv.acc_tot[chain];
    const int nsteps = ra.mode == 1 ? 1 : ra.nsteps;
    const long long grp = sv.cpg > 0 ? gchain / sv.cpg : 0;

    uint4 sel = chain_rng(sv.seed, gchain, ra.step0, RNG_SELECT);
    bool use_de, jump;
    double dz[ND];
    {
        const DeSel de = de_select(sv, ra, chain, gchain, ra.step0, sel);
        use_de = de.use; jump = de.jump;
#pragma unroll
        for (int d = 0; d < ND; d++) dz[d] = use_de ? sv.hist[de.o1 + d] - sv.hist[de.o2 + d] : 0.0;
    }
    for (int s = 0; s < nsteps; s++) {
        const long long i = ra.step0 + s;
        if (ra.mode == 0 && ra.tune && i > 0 && (i % sv.tune_interval) == 0) {
            double f = tune_factor((double)acc_win / (double)sv.tune_interval);
            if (sv.tune_target == 0) lamb *= f; else scal *= f;
            acc_win = 0;
        }
        // ---- next step's selection; its rows start their way from HBM now
        const bool more = ra.mode == 0 && s + 1 < nsteps;
        uint4 sel_n = sel;
        DeSel dn{0, 0, false, false};
        bool hz1 = false, hz2 = false;
        if (more) {
            sel_n = chain_rng(sv.seed, gchain, i + 1, RNG_SELECT);
            dn = de_select(sv, ra, chain, gchain, i + 1, sel_n);
        }
        if (dn.use) {
            const size_t hcur = hist_off(sv, i, chain);
            hz1 = dn.o1 == hcur; hz2 = dn.o2 == hcur;     // a row this step is about to write: taken from registers below
#pragma unroll
            for (int d = 0; d < ND; d++) {
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(z1a + d * 512u), "l"(sv.hist + dn.o1 + d) : "memory");
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(z2a + d * 512u), "l"(sv.hist + dn.o2 + d) : "memory");
            }
        }
        // ---- proposal, transforms, priors
        // the value registers the CE model reads: 0 log10 Ae, 1 log10 F, 2 E, 3 m, 4 b, VAL_NU (every dimension overwrites
        // its register in every step, so starting from the defaults each time is the same as carrying them)
        double val[6];
#pragma unroll
        for (int k = 0; k < 5; k++) val[k] = sv.val_defaults[k];
        val[5] = sv.val_defaults[VAL_NU];
        double lps = 0.0;
#pragma unroll
        for (int d = 0; d < ND; d++) {
            double qn = qs[d];
            if (ra.mode == 0) {
                double eps = proposal_noise(sv.proposal_dist, chain_rng(sv.seed, gchain, i, (unsigned int)d));
                if (use_de) qn += (jump ? 1.0 : lamb) * dz[d];
                qn += eps * scal;
            }
            qp[d * 64] = qn;
            double v;
            lps += transform_dim(sv.pr[d], qn, v);
            vp[d * 64] = v;
            const int tg = sv.pr[d].target;
#pragma unroll
            for (int k = 0; k < 5; k++) val[k] = tg == k ? v : val[k];
            val[5] = tg == VAL_NU ? v : val[5];
        }
        // ---- forward model + likelihood: as in k_mcmc_ng
        double Ae = exp10(val[0]), F = exp10(val[1]), E = val[2];
        double T = (E - val[4]) / val[3];
        double P = ce_lapse_rate_step(E);
        const CeStep cs = ce_step(T, P);
        double nu = sv.nu_sampled ? sv.nu_lo + (sv.nu_hi - sv.nu_lo) * val[5] : sv.nu_fixed;
        double cst = sv.lik_kind == 1 ? lik_studentt_const(nu) : 0.0;
        double ll = 0.0;
        for (int g = 0; g < sv.gases.n; g++) {
            double mu = ce_exc_step(sv.gases.id[g], cs, E, T, Ae, F, P);
            double ob = sv.cpg > 0 ? sv.g_obs[grp * sv.gases.n + g] : sv.obs[g];
            double is = sv.cpg > 0 ? sv.g_isd[grp * sv.gases.n + g] : sv.isd[g];
            double lc = sv.cpg > 0 ? sv.g_lc[grp * sv.gases.n + g] : sv.lc[g];
            ll += sv.lik_kind == 1 ? lik_term_studentt(ob, mu, is, lc, nu, cst) : lik_term_normal(ob, mu, is, lc);
        }
        double lpn = lps + ll;
        double delta = lpn - logp;
        bool acc = ra.mode == 1 || (isfinite(delta) && log(u01(sel.z, sel.w)) < delta);
        if (acc) {
#pragma unroll
            for (int d = 0; d < ND; d++) { qs[d] = qp[d * 64]; vn[d * 64] = vp[d * 64]; }
            logp = lpn;
            if (ra.mode == 0) { acc_win++; acc_tot++; }
        }
        if (ra.mode == 0) {
            size_t ho = hist_off(sv, i, chain);
#pragma unroll
            for (int d = 0; d < ND; d++) sv.hist[ho + d] = qs[d];
            if (ra.record && ((i - ra.step0) % ra.thin) == 0) {
                long long draw = ra.draw0 + (i - ra.step0) / ra.thin;
#pragma unroll
                for (int d = 0; d < ND; d++) {          // record_dim() on the shared-memory copy
                    const double v = vn[d * 64];
                    if (ra.trace) ra.trace[((size_t)(draw - ra.draw0) * (size_t)sv.B + (size_t)chain) * ND + d] = v;
                    const double n = (double)(draw + 1);
                    double mean = wmean[d * 64];
                    const double dl = v - mean;
                    mean += dl / n;
                    wmean[d * 64] = mean;
                    wm2[d * 64] += dl * (v - mean);
                }
            }
        }
        // ---- hand the prefetched rows to the next step
        sel = sel_n;
        use_de = dn.use; jump = dn.jump;
        asm volatile("cp.async.wait_all;" ::: "memory");
        if (use_de) {
#pragma unroll
            for (int d = 0; d < ND; d++) dz[d] = (hz1 ? qs[d] : z1s[d * 64]) - (hz2 ? qs[d] : z2s[d * 64]);
        }
    }
#pragma unroll
    for (int d = 0; d < ND; d++) sv.q[chain * ND + d] = qs[d];
    if (rec) {
#pragma unroll
        for (int d = 0; d < ND; d++) { sv.wf_mean[chain * ND + d] = wmean[d * 64]; sv.wf_m2[chain * ND + d] = wm2[d * 64]; }
    }
    sv.logp[chain] = logp;
    sv.lamb[chain] = lamb;
    sv.scal[chain] = scal;
    sv.acc_win[chain] = acc_win;
    sv.acc_tot[chain] = acc_tot;
}

#ifndef NGRTD_NO_AUX_KERNELS   // compiled into part 0 only (ngrtd_api.cu, build partitioning)
__global__ void k_philox_kat(uint4 c, uint2 k, unsigned int* out) {
    uint4 r = philox4x32_10(c, k);
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}
#endif

}  // namespace ngrtd
