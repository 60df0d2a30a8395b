/*
 * TEST INFRASTRUCTURE ONLY -- plain-C restatement of the reference hot path (one chain at a time, the
 * reference's own per-lag arithmetic: a true exp() per lag, two-pass normalisation, a dot product).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity status: PINNED against the golden vectors generated from the untouched reference
 * (tests/test_oracle_golden.py::test_c_oracle_*).  Citations are relative to /root/reference.
 *
 * Build: make -C oracle   (gcc -O2 -pthread -shared; no -ffast-math: the oracle must keep IEEE semantics)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

enum { MOD_NONE = 0, MOD_PISTON = 1, MOD_EXPONENTIAL = 2, MOD_EPM = 3, MOD_DISPERSION = 4 };
enum { P_TAU1 = 0, P_TAU2, P_F1, P_F2, P_ETA1, P_ETA2, P_D1, P_D2, P_J, P_THALF_CFC, P_LAMSF6, NSLOT };
enum { ACC_NONE = 0, ACC_3HE = 1, ACC_4HE = 2 };

typedef struct {
    int32_t series;        /* column of the series matrix or -1 (zeros) */
    int32_t rad_accum;
    double lambda;
    int32_t use_thalf_cfc;
    int32_t use_lamsf6;
} oracle_tracer;

static double j_flux(double Del, double rho_r, double rho_w, double U, double Th, double phi) {
    const double PU = 1.19e-13, PTh = 2.88e-14;                 /* utils/noble_gas_utils.py:335-348 */
    return Del * rho_r / rho_w * (U * PU + Th * PTh) * ((1 - phi) / phi);
}

/* utils/convolution_integral_utils.py:168-196,270 -- normalised weights of one model into g[L] */
static void gen_g_tp(int mod, int L, const double* tp, double tau, double eta, double D, double* g) {
    int k;
    if (mod == MOD_PISTON) {                                    /* :178-181, first index wins ties */
        int ix = 0;
        double best = fabs(tp[0] - tau);
        for (k = 1; k < L; k++) {
            double d = fabs(tp[k] - tau);
            if (d < best) { best = d; ix = k; }
        }
        if (best != best) ix = 0;                               /* numpy argmin over NaN -> 0 */
        for (k = 0; k < L; k++) g[k] = 0.0;
        g[ix] = 1.0;
    } else if (mod == MOD_EXPONENTIAL) {                        /* :184 */
        for (k = 0; k < L; k++) g[k] = (1. / tau) * exp(-tp[k] / tau);
    } else if (mod == MOD_EPM) {                                /* :187-190 */
        double thr = tau * (1 - (1 / eta));
        for (k = 0; k < L; k++) g[k] = (tp[k] >= thr) ? (eta / tau) * exp(-(eta * tp[k] / tau) + eta - 1.) : 0.0;
    } else {                                                    /* :194-196 */
        for (k = 0; k < L; k++) {
            double x = tp[k] / tau;
            double f1 = (1. / tau) / (sqrt(4. * M_PI * D * x));
            double f2 = (1. / x) * exp(-1. * (((1. - x) * (1. - x)) / (4. * D * x)));
            g[k] = f1 * f2;
        }
    }
    {                                                           /* :270 */
        double S = 0.0;
        for (k = 0; k < L; k++) S += g[k];
        for (k = 0; k < L; k++) g[k] = g[k] / S;
    }
}

/* :300-340 for one tracer given normalised g */
static double convolve(int L, const double* tp, const double* g, const double* c, int cstride, const double* idx,
                       double lam, int rad_accum, double J) {
    double acc = 0.0;
    int k;
    for (k = 0; k < L; k++) {
        double dec = exp(-lam * tp[k]);
        double gd = g[k] * (rad_accum == ACC_3HE ? (1 - dec) : dec);     /* :313-316 */
        double ck = c ? c[(size_t)k * cstride] : 0.0;
        if (rad_accum == ACC_4HE) ck = ck + (idx ? idx[k] : (double)k) * J;  /* :320-327 */
        acc += ck * gd;                                                   /* :336-337 */
    }
    return acc;
}

/* Batched ForwardMod.perform (age_ens_runs_mcmc/run_age_mcmc_utils.py:81-163) for all tracers of a chain.
 * The RTD of a chain is generated once and shared by its tracers (the reference regenerates it per tracer
 * with identical arithmetic, so results are the same).  Chains are split over POSIX threads. */
typedef struct {
    int32_t L, nseries, ntracer, mod1, mod2, ndim;
    const double *series, *lag_index, *theta, *tp;
    const oracle_tracer* tr;
    const int* col_of_slot;
    double* out;
    int64_t b0, b1;
} fwd_job;

static void* fwd_worker(void* arg) {
    const fwd_job* j = (const fwd_job*)arg;
    const int L = j->L;
    double* g1 = (double*)malloc(sizeof(double) * L);
    double* g2 = (double*)malloc(sizeof(double) * L);
    int64_t b;
    for (b = j->b0; b < j->b1; b++) {
        const double* row = j->theta + b * j->ndim;
        double p[NSLOT];
        int has[NSLOT], t;
        for (t = 0; t < NSLOT; t++) { has[t] = j->col_of_slot[t] >= 0; p[t] = has[t] ? row[j->col_of_slot[t]] : 0.0; }
        if (!has[P_F1]) p[P_F1] = 1.0;                       /* p_dict defaults :73-79 */
        if (j->mod1 == MOD_EXPONENTIAL) p[P_ETA1] = 1.0;
        if (j->mod2 == MOD_EXPONENTIAL) p[P_ETA2] = 1.0;
        gen_g_tp(j->mod1, L, j->tp, p[P_TAU1], p[P_ETA1], p[P_D1], g1);
        if (j->mod2 != MOD_NONE) gen_g_tp(j->mod2, L, j->tp, p[P_TAU2], p[P_ETA2], p[P_D2], g2);
        for (t = 0; t < j->ntracer; t++) {
            const oracle_tracer* tr = j->tr + t;
            double lam = tr->lambda, J = 0.0, c1, c2 = 0.0, cout;
            const double* c = tr->series >= 0 ? j->series + tr->series : NULL;
            if (tr->rad_accum == ACC_4HE) {                  /* :90-91, :101 */
                double lj = has[P_J] ? p[P_J] : log10(j_flux(1., 2700, 1000, 3.0, 10.0, 0.05));
                J = pow(10.0, lj);
            }
            if (tr->use_thalf_cfc && has[P_THALF_CFC]) lam = -1 * log(0.5) / p[P_THALF_CFC];   /* :107-109 */
            c1 = convolve(L, j->tp, g1, c, j->nseries, j->lag_index, lam, tr->rad_accum, J);
            if (j->mod2 != MOD_NONE) c2 = convolve(L, j->tp, g2, c, j->nseries, j->lag_index, lam, tr->rad_accum, J);
            cout = p[P_F1] * c1 + p[P_F2] * c2;              /* :154 */
            if (tr->use_lamsf6) cout *= (1 + p[P_LAMSF6]);   /* :160-161 */
            j->out[b * j->ntracer + t] = cout;
        }
    }
    free(g1);
    free(g2);
    return NULL;
}

int oracle_max_threads(void) {
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

int oracle_forward(int32_t L, int32_t nseries, const double* series, const double* lag_index, double dtp,
                   int32_t ntracer, const oracle_tracer* tr, int32_t mod1, int32_t mod2, const double* theta,
                   int64_t B, int32_t ndim, const int32_t* slot_of_col, double* out, int32_t nthreads) {
    int col_of_slot[NSLOT];
    int s, i, nt;
    double* tp = (double*)malloc(sizeof(double) * L);
    fwd_job* jobs;
    pthread_t* th;
    for (s = 0; s < NSLOT; s++) col_of_slot[s] = -1;
    for (i = 0; i < ndim; i++) col_of_slot[slot_of_col[i]] = i;
    for (i = 0; i < L; i++) tp[i] = (double)i;
    tp[0] += 1e-5;                                               /* conv utils :168-173 */
    for (i = 0; i < L; i++) tp[i] += dtp;
    nt = nthreads > 0 ? nthreads : oracle_max_threads();
    if ((int64_t)nt > B) nt = B > 0 ? (int)B : 1;
    jobs = (fwd_job*)malloc(sizeof(fwd_job) * nt);
    th = (pthread_t*)malloc(sizeof(pthread_t) * nt);
    for (i = 0; i < nt; i++) {
        fwd_job j = {L, nseries, ntracer, mod1, mod2, ndim, series, lag_index, theta, tp, tr, col_of_slot, out,
                     B * i / nt, B * (i + 1) / nt};
        jobs[i] = j;
        if (i > 0) pthread_create(&th[i], NULL, fwd_worker, &jobs[i]);
    }
    fwd_worker(&jobs[0]);
    for (i = 1; i < nt; i++) pthread_join(th[i], NULL);
    free(jobs);
    free(th);
    free(tp);
    return nt;
}

/* ---------------- closed-equilibrium model, utils/noble_gas_utils.py:103-253 ---------------- */
static const double ATM_STD[5] = {5.24e-6, 1.818e-5, 9.34e-3, 1.14e-6, 8.7e-8};
static const double SOL[5][4] = {{-0.00953, 0.107722, 0.001969, -0.043825}, {-7.259, 6.95, -1.3826, 0.0538},
                                 {-9.52, 8.83, -1.8959, 0.0698}, {-6.292, 5.612, -0.8881, -0.0458},
                                 {-3.902, 2.439, 0.3863, -0.221}};
static const double SETCH[5][3] = {{-10.081, 15.1068, 4.8127}, {-11.9556, 18.4062, 5.5464}, {-10.6951, 16.7513, 4.9551},
                                   {-9.9787, 15.7619, 4.6181}, {-14.5524, 22.5255, 6.7513}};

static double poly(const double* A, double T_k) {
    double t = .001 * T_k;
    return A[0] + (A[1] / t) + (A[2] / (t * t)) + (A[3] / (t * t * t));
}
static double solubility(int gas, double T, double S) {          /* :117-180 */
    double T_k = T + 273.15, gamma = 1.0, K_h;
    if (T < 65.) gamma = exp(S * (SETCH[gas][0] + (SETCH[gas][1] / (.01 * T_k)) + (SETCH[gas][2] * log(.01 * T_k))));
    if (gas == 0) {
        double F = exp(poly(SOL[0], T_k));
        double X_Ar_water = 1. / (exp(poly(SOL[2], T_k))) * 9.31e-3;
        K_h = 5.24e-6 / (F * (5.24e-6 / 9.31e-3) * X_Ar_water);
    } else {
        K_h = exp(poly(SOL[gas], T_k));
    }
    return gamma * K_h;
}
static double vapor_pressure(double T) {                          /* :184-199 */
    double A = T <= 99.0 ? 8.07131 : 8.14019, Bc = T <= 99.0 ? 1730.63 : 1810.94, C = T <= 99.0 ? 233.426 : 244.485;
    double P = pow(10.0, A - (Bc / (C + T)));
    P = P / 760. * 101325;
    return P / 1.0e9;
}
/* what: 0 ce_exc(True), 1 ce_exc(False), 2 equil_conc_dry, 3 equil_conc, 4 solubility; P NULL = lapse rate */
int oracle_ce(int32_t what, int32_t ngas, const int32_t* gases, const double* E, const double* T, const double* Ae,
              const double* F, const double* P, double S, int64_t B, double* out) {
    int64_t i;
    for (i = 0; i < B; i++) {
        double Pi = P ? P[i] : pow(1 - .0065 * E[i] / 288.15, 5.2561) * 0.000101325;   /* :112 */
        double pv = vapor_pressure(T[i]);
        int g;
        for (g = 0; g < ngas; g++) {
            int gas = gases[g];
            double K = solubility(gas, T[i], S), z = ATM_STD[gas], v;
            if (what == 4) v = K;
            else if (what == 3) v = (z * Pi / K) * (22414. / 18.);
            else {
                double C_eq = (T[i] < 0.0) ? -9999.0 : ((z * (Pi - pv)) / K) * (22414. / 18.);   /* :225-230 */
                if (what == 2) v = C_eq;
                else {
                    double C_ex = ((1 - F[i]) * Ae[i] * z) / (1 + ((F[i] * Ae[i] * z) / C_eq));     /* :248 */
                    v = what == 0 ? C_ex + C_eq : C_ex;
                }
            }
            out[i * ngas + g] = v;
        }
    }
    return 0;
}

/* ---------------- likelihoods (pymc3 3.11.2 formulae) ---------------- */
int oracle_loglik(int32_t kind, int32_t T, const double* mu, const double* obs, const double* sd, const double* nu,
                  int64_t B, double* logp) {
    int64_t i;
    for (i = 0; i < B; i++) {
        double acc = 0.0;
        int t;
        for (t = 0; t < T; t++) {
            double m = mu[i * T + t];
            if (kind == 0) {
                acc += -0.5 * log(2 * M_PI * sd[t] * sd[t]) - (obs[t] - m) * (obs[t] - m) / (2 * sd[t] * sd[t]);
            } else {
                double n = nu[i], lam = 1.0 / (sd[t] * sd[t]);
                acc += lgamma((n + 1.0) / 2.0) - lgamma(n / 2.0) + 0.5 * log(lam / (n * M_PI)) -
                       (n + 1.0) / 2.0 * log1p(lam * (obs[t] - m) * (obs[t] - m) / n);
            }
        }
        logp[i] = acc;
    }
    return 0;
}

