// ngrtd_api.cu -- C ABI of libngrtd.so (include/ngrtd.h): plan management, kernel dispatch, host-buffer variants.
// No torch types anywhere; device pointers and a stream handle come from the caller.
#include "../../include/ngrtd.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <string>
#include <vector>

// Build partitioning (build time only): the file compiles as ONE translation unit by default; __graft_entry__.build() compiles
// it three times in parallel with -DNGRTD_PART=0 (C ABI, small kernels), 1 (the k_forward instantiations) and 2 (the
// k_mcmc_age instantiations) and links the objects -- same source for every kernel, a third of the wall time.
#if !defined(NGRTD_PART)
#define NGRTD_HAS_API 1
#define NGRTD_HAS_FWD 1
#define NGRTD_HAS_MCMC 1
#elif NGRTD_PART == 0
#define NGRTD_HAS_API 1
#define NGRTD_HAS_FWD 0
#define NGRTD_HAS_MCMC 0
#elif NGRTD_PART == 1
#define NGRTD_HAS_API 0
#define NGRTD_HAS_FWD 1
#define NGRTD_HAS_MCMC 0
#else
#define NGRTD_HAS_API 0
#define NGRTD_HAS_FWD 0
#define NGRTD_HAS_MCMC 1
#endif
#if !NGRTD_HAS_API
#define NGRTD_NO_AUX_KERNELS 1      // the non-template kernels of the headers (k_ce, k_cfc, k_mcmc_ng, ...) live in part 0 only
#endif

#include "ngrtd_common.cuh"
#include "ngrtd_forward.cuh"
#include "ngrtd_ce.cuh"
#include "ngrtd_mcmc.cuh"

using namespace ngrtd;

#if NGRTD_HAS_API
thread_local std::string ngrtd_g_err;          // one object per process: the parts of a partitioned build share it
#else
extern thread_local std::string ngrtd_g_err;
#endif
#define g_err ngrtd_g_err

static int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CUDA_TRY(x)                                                                                   \
    do {                                                                                              \
        cudaError_t e__ = (x);                                                                        \
        if (e__ != cudaSuccess)                                                                       \
            return fail(NGRTD_ECUDA, std::string(#x) + ": " + cudaGetErrorString(e__));               \
    } while (0)
// same, releasing a half-built object first (create paths)
#define CUDA_TRY_OR(x, cleanup)                                                                       \
    do {                                                                                              \
        cudaError_t e__ = (x);                                                                        \
        if (e__ != cudaSuccess) {                                                                     \
            cleanup;                                                                                  \
            return fail(NGRTD_ECUDA, std::string(#x) + ": " + cudaGetErrorString(e__));               \
        }                                                                                             \
    } while (0)

// Entry points that launch or copy run on the device their plan / sampler lives on and leave the caller's current
// device as they found it (one thread may drive objects on several GPUs).
struct DeviceGuard {
    int prev = -1;
    bool good = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { prev = -1; good = false; return; }
        if (prev != dev) good = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard() {
        int cur = -1;
        if (prev >= 0 && cudaGetDevice(&cur) == cudaSuccess && cur != prev) cudaSetDevice(prev);
    }
    bool ok() const { return good; }
};

constexpr int HOST_PARTS_MAX = 8;
#ifndef NGRTD_STAGE_DEFAULT
#define NGRTD_STAGE_DEFAULT 0
#endif

struct ngrtd_plan {
    int device = 0;
    int nsm = 0;
    int L = 0, Lpad = 0;
    int mod1 = 0, mod2 = 0;
    int cls1 = 0, cls2 = 0;
    bool dyn = false;
    PlanView pv{};
    double *dXf = nullptr, *dXd = nullptr, *ditp = nullptr, *dxraw = nullptr, *dxrawd = nullptr, *dtbl = nullptr, *dtbl11 = nullptr;
    // workspace of the *_host entry points
    double *w_theta = nullptr, *w_out = nullptr, *w_logp = nullptr, *w_nu = nullptr;
    size_t w_theta_n = 0, w_out_n = 0, w_logp_n = 0, w_nu_n = 0;
    // host-buffer pipeline: copy-in stream, compute stream, copy-out stream + one event pair per part
    cudaStream_t hstream = nullptr, hstream2 = nullptr, hstream3 = nullptr;
    cudaEvent_t ev_in[HOST_PARTS_MAX] = {}, ev_k[HOST_PARTS_MAX] = {};
    // submit / wait slots (ngrtd_forward_loglik_host_submit): own device buffers and events per slot
    struct HostSlot {
        double *theta = nullptr, *nu = nullptr, *logp = nullptr, *out = nullptr;
        size_t theta_n = 0, nu_n = 0, logp_n = 0, out_n = 0;
        cudaEvent_t ev_in = nullptr, ev_k = nullptr, ev_done = nullptr;
        cudaStream_t s_k = nullptr;     // the slot's own kernel stream: kernels of different slots are NOT stream-ordered, so the
                                        // CTAs of batch i+1 move onto SMs as the CTAs of batch i retire (no tail / head gap)
        bool busy = false;
    } slots[NGRTD_HOST_SLOTS];
};

static int cls_of(int mod) {
    switch (mod) {
        case NGRTD_MOD_NONE: return CLS_NONE;
        case NGRTD_MOD_PISTON: return CLS_P;
        case NGRTD_MOD_EXPONENTIAL:
        case NGRTD_MOD_EXP_PIST_FLOW: return CLS_G;
        case NGRTD_MOD_DISPERSION: return CLS_D;
        default: return -1;
    }
}

#if NGRTD_HAS_API    // ---- part 0: C ABI (plans)
extern "C" int ngrtd_version(void) { return NGRTD_VERSION; }
extern "C" int ngrtd_host_alloc(void** out, size_t bytes, int32_t write_combined) {
    if (!out || bytes == 0) return fail(NGRTD_EINVAL, "host_alloc: null pointer or zero size");
    *out = nullptr;
    CUDA_TRY(cudaHostAlloc(out, bytes, cudaHostAllocPortable | (write_combined ? cudaHostAllocWriteCombined : 0)));
    return NGRTD_OK;
}
extern "C" int ngrtd_host_free(void* p) {
    if (p) CUDA_TRY(cudaFreeHost(p));
    return NGRTD_OK;
}
extern "C" int ngrtd_build_features(void) {
    int f = 0;
    if (DM_TAIL) f |= NGRTD_FEATURE_DM_TAIL;
#ifdef NGRTD_XF_SWIZZLE
    f |= NGRTD_FEATURE_XF_SWIZZLE;
#endif
#ifdef NGRTD_TP_DADD
    f |= NGRTD_FEATURE_TP_DADD;
#endif
    return f;
}
extern "C" const char* ngrtd_last_error(void) { return g_err.c_str(); }

static double j_flux(double Del, double rho_r, double rho_w, double U, double Th, double phi) {
    const double PU = 1.19e-13, PTh = 2.88e-14;   // utils/noble_gas_utils.py:335-348
    return Del * rho_r / rho_w * (U * PU + Th * PTh) * ((1 - phi) / phi);
}

extern "C" int ngrtd_plan_create(ngrtd_plan** out, int32_t L, int32_t nseries, const double* series,
                                 const double* lag_index, double dtp, int32_t ntracer, const ngrtd_tracer* tracers,
                                 int32_t mod_type1, int32_t mod_type2, int32_t device) {
    if (!out) return fail(NGRTD_EINVAL, "plan: null output pointer");
    *out = nullptr;
    if (L < 1) return fail(NGRTD_EINVAL, "plan: L must be >= 1");
    if (nseries < 0 || (nseries > 0 && !series)) return fail(NGRTD_EINVAL, "plan: series is null");
    if (ntracer < 1 || ntracer > MAX_TRACER) return fail(NGRTD_EINVAL, "plan: ntracer must be in 1..8");
    if (!tracers) return fail(NGRTD_EINVAL, "plan: tracers is null");
    int c1 = cls_of(mod_type1), c2 = cls_of(mod_type2);
    if (c1 <= 0) return fail(NGRTD_EINVAL, "plan: unknown mod_type1 " + std::to_string(mod_type1));
    if (c2 < 0) return fail(NGRTD_EINVAL, "plan: unknown mod_type2 " + std::to_string(mod_type2));
    if (dtp != std::floor(dtp)) return fail(NGRTD_EINVAL, "plan: dtp must be integer valued (np.floor in the reference)");
    int dev = device;
    if (dev < 0) CUDA_TRY(cudaGetDevice(&dev));
    DeviceGuard guard(dev);                 // the caller's current device is restored on return
    if (!guard.ok()) return fail(NGRTD_ECUDA, "plan: cudaSetDevice(" + std::to_string(dev) + ") failed");

    auto* P = new ngrtd_plan();
    P->device = dev;
    CUDA_TRY_OR(cudaDeviceGetAttribute(&P->nsm, cudaDevAttrMultiProcessorCount, dev), delete P);
    P->L = L;
    P->Lpad = (L + 3) & ~3;
    P->mod1 = mod_type1;
    P->mod2 = mod_type2;
    P->cls1 = c1;
    P->cls2 = c2;
    const int Lpad = P->Lpad;

    // lag grid exactly as the reference builds it: arange -> tp[0] += 1e-5 -> tp += dtp (conv utils :168-173)
    std::vector<double> tp(Lpad, 0.0), p15(Lpad, 0.0), itp(2 * (size_t)Lpad, 0.0);   // itp: {1/tp, tp} pairs
    for (int k = 0; k < L; k++) {
        double t = (double)k;
        if (k == 0) t += 1e-5;
        t += dtp;
        tp[k] = t;
        itp[2 * k] = 1.0 / t;
        itp[2 * k + 1] = t;
        p15[k] = 1.0 / (t * std::sqrt(t));
    }
    // folded columns
    struct ColKey { int series; int mode; double lambda; };   // mode 0 decay, 1 ingrowth, 2 lag-index*decay
    std::vector<ColKey> cols;
    cols.push_back({-2, 0, 0.0});   // column 0: ones
    auto find_col = [&](int s, int mode, double lam) -> int {
        for (size_t i = 1; i < cols.size(); i++)
            if (cols[i].series == s && cols[i].mode == mode && cols[i].lambda == lam) return (int)i;
        cols.push_back({s, mode, lam});
        return (int)cols.size() - 1;
    };
    int ndyn = 0, dyn_series = -1;
    for (int t = 0; t < ntracer; t++) {
        const ngrtd_tracer& tr = tracers[t];
        if (tr.series >= nseries) { delete P; return fail(NGRTD_EINVAL, "plan: tracer series index out of range"); }
        TracerDev td{-1, -1, 0, tr.use_lamsf6 ? 1 : 0};
        bool zero_series = tr.series < 0;
        if (!zero_series) {
            zero_series = true;
            for (int k = 0; k < L && zero_series; k++) zero_series = (series[(size_t)k * nseries + tr.series] == 0.0);
        }
        if (tr.use_thalf_cfc) {
            if (tr.rad_accum != NGRTD_ACC_NONE) { delete P; return fail(NGRTD_EINVAL, "plan: use_thalf_cfc with rad_accum is not supported"); }
            if (ndyn && dyn_series != tr.series) { delete P; return fail(NGRTD_EINVAL, "plan: only one per-chain-lambda series per plan"); }
            ndyn++;
            dyn_series = tr.series;
            td.dyn = 1;
        } else if (tr.rad_accum == NGRTD_ACC_3HE) {
            if (!zero_series) td.col_a = find_col(tr.series, 1, tr.lambda);
        } else if (tr.rad_accum == NGRTD_ACC_4HE) {
            if (!zero_series) td.col_a = find_col(tr.series, 0, tr.lambda);
            td.col_b = find_col(-1, 2, tr.lambda);
        } else if (tr.rad_accum == NGRTD_ACC_NONE) {
            if (!zero_series) td.col_a = find_col(tr.series, 0, tr.lambda);
        } else {
            delete P;
            return fail(NGRTD_EINVAL, "plan: unknown rad_accum");
        }
        P->pv.tr[t] = td;
    }
    if ((int)cols.size() > NCOL) {
        delete P;
        return fail(NGRTD_EINVAL, "plan: more than 7 distinct folded columns; split the tracers over two plans");
    }
    P->dyn = ndyn > 0;
    std::vector<double> Xf((size_t)Lpad * NCOL, 0.0), Xd((size_t)Lpad * NCOL, 0.0), xraw(Lpad, 0.0), xrawd(Lpad, 0.0);
    for (int k = 0; k < L; k++) {
        for (size_t c = 0; c < cols.size(); c++) {
            double v;
            const ColKey& ck = cols[c];
            if (c == 0) {
                v = 1.0;
            } else {
                double dec = std::exp(-ck.lambda * tp[k]);                       // conv utils :316
                if (ck.mode == 1) v = series[(size_t)k * nseries + ck.series] * (1 - dec);   // :314
                else if (ck.mode == 2) v = (lag_index ? lag_index[k] : (double)k) * dec;     // :323 index * J
                else v = series[(size_t)k * nseries + ck.series] * dec;
            }
            Xf[(size_t)xf_index(k, (int)c)] = v;
            Xd[(size_t)xf_index(k, (int)c)] = v * p15[k];
        }
        if (P->dyn && dyn_series >= 0) {
            xraw[k] = series[(size_t)k * nseries + dyn_series];
            xrawd[k] = xraw[k] * p15[k];
        }
    }
    // ---- constant-tail detection (see PlanView::Kc) ----
    {
        PlanView& pv = P->pv;
        pv.Kc = Lpad;
        pv.dyn_bg = 0.0;
        for (int c = 0; c < NCOL; c++) pv.ct[c] = ColTail{-1, 0.0, 0.0, 0.0, 0.0};
        const bool any_d = c1 == CLS_D || c2 == CLS_D, any_g = c1 == CLS_G || c2 == CLS_G;
        // dispersion components: by quadrature (DM_TAIL builds, the default)
        const bool want = getenv("NGRTD_NO_TAIL") == nullptr && (any_d ? DM_TAIL : any_g);
        if (want && L >= 64) {
            int kvar = 0;                                       // first lag from which every used series is constant
            auto scan = [&](int sidx) {
                const double last = series[(size_t)(L - 1) * nseries + sidx];
                int k = L - 1;
                while (k > 0 && series[(size_t)(k - 1) * nseries + sidx] == last) k--;
                kvar = std::max(kvar, k);
            };
            bool ok = true;
            double i0 = 0.0, sl = 1.0;
            for (size_t c = 1; c < cols.size(); c++) {
                if (cols[c].mode == 2) {                        // lag index must be an exact arithmetic progression in the tail
                    if (lag_index) {
                        sl = lag_index[L - 1] - lag_index[L - 2];
                        i0 = lag_index[L - 1] - sl * (double)(L - 1);
                        int k = L - 1;
                        while (k > 0 && lag_index[k - 1] == i0 + sl * (double)(k - 1)) k--;
                        kvar = std::max(kvar, k);
                    }
                } else {
                    scan(cols[c].series);
                }
            }
            if (P->dyn && dyn_series >= 0) scan(dyn_series);
            int Kc = std::max(4, (kvar + 3) & ~3);
            // dispersion tails are integrated by quadrature with Euler-Maclaurin end terms, validated (profiles/
            // r1_dm_tail_quadrature_study.txt) for a cut at lag >= 128: closer to lag 0 the weight t^-1.5 exp(-a/t - c t)
            // varies too fast across one lag for the end correction (2e-8 at Kc = 4, tau = 37.5, D = 0.3)
            if (any_d) Kc = std::max(Kc, 128);
            if (ok && Kc + 32 <= L) {                           // a tail worth cutting
                pv.Kc = Kc;
                pv.ct[0] = ColTail{0, 1.0, 0.0, 0.0, 0.0};
                for (size_t c = 1; c < cols.size(); c++) {
                    const ColKey& ck = cols[c];
                    if (ck.mode == 2) pv.ct[c] = ColTail{3, 0.0, ck.lambda, i0, sl};
                    else pv.ct[c] = ColTail{ck.mode == 1 ? 2 : 1, series[(size_t)(L - 1) * nseries + ck.series], ck.lambda, 0.0, 0.0};
                }
                if (P->dyn && dyn_series >= 0) pv.dyn_bg = series[(size_t)(L - 1) * nseries + dyn_series];
            }
        }
    }
    // exp tables of exp_scaled_bits<TB>: 2^(i/N) with i << (20 - log2 N) and exp_tbl_fold<TB>() subtracted from the high word,
    // so that the exponent insertion becomes one integer multiply-add; the kernels copy (and for N = 128 replicate) them
    // into shared memory
    auto make_tbl = [](int bits, int sub, unsigned int fold) {
        const int n = 1 << bits;
        std::vector<double> tbl(n);
        for (int i = 0; i < n; i++) {
            double v = std::exp2((double)i / n);
            uint64_t b;
            std::memcpy(&b, &v, 8);
            uint32_t hi = (uint32_t)(b >> 32) - ((uint32_t)(i >> sub) << (20 - (bits - sub))) - fold;   // modulo 2^32, like the device add
            b = ((uint64_t)hi << 32) | (b & 0xffffffffull);
            std::memcpy(&tbl[i], &b, 8);
        }
        return tbl;
    };
    const std::vector<double> tbl7 = make_tbl(7, ExpCfg<7>::SUB, exp_tbl_fold<7>()), tbl11 = make_tbl(11, ExpCfg<11>::SUB, exp_tbl_fold<11>());

    auto up = [&](double** d, const std::vector<double>& h) -> cudaError_t {
        cudaError_t e = cudaMalloc((void**)d, h.size() * sizeof(double));
        if (e != cudaSuccess) return e;
        return cudaMemcpy(*d, h.data(), h.size() * sizeof(double), cudaMemcpyHostToDevice);
    };
    cudaError_t e = cudaSuccess;
    if ((e = up(&P->dXf, Xf)) != cudaSuccess || (e = up(&P->dXd, Xd)) != cudaSuccess ||
        (e = up(&P->ditp, itp)) != cudaSuccess || (e = up(&P->dxraw, xraw)) != cudaSuccess ||
        (e = up(&P->dxrawd, xrawd)) != cudaSuccess || (e = up(&P->dtbl, tbl7)) != cudaSuccess ||
        (e = up(&P->dtbl11, tbl11)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&P->hstream, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&P->hstream2, cudaStreamNonBlocking)) != cudaSuccess ||
        (e = cudaStreamCreateWithFlags(&P->hstream3, cudaStreamNonBlocking)) != cudaSuccess) {
        ngrtd_plan_destroy(P);
        return fail(NGRTD_ECUDA, std::string("plan upload: ") + cudaGetErrorString(e));
    }
    for (int i = 0; i < HOST_PARTS_MAX && e == cudaSuccess; i++) {
        e = cudaEventCreateWithFlags(&P->ev_in[i], cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&P->ev_k[i], cudaEventDisableTiming);
    }
    for (int i = 0; i < NGRTD_HOST_SLOTS && e == cudaSuccess; i++) {
        e = cudaEventCreateWithFlags(&P->slots[i].ev_in, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&P->slots[i].ev_k, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaEventCreateWithFlags(&P->slots[i].ev_done, cudaEventDisableTiming);
        if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&P->slots[i].s_k, cudaStreamNonBlocking);
    }
    if (e != cudaSuccess) {
        ngrtd_plan_destroy(P);
        return fail(NGRTD_ECUDA, std::string("plan upload: ") + cudaGetErrorString(e));
    }
    PlanView& pv = P->pv;
    pv.L = L;
    pv.Lpad = Lpad;
    pv.dtp = dtp;
    pv.Xf = P->dXf;
    pv.Xd = P->dXd;
    pv.itp = P->ditp;
    pv.xraw = P->dxraw;
    pv.xrawd = P->dxrawd;
    pv.tbl7 = P->dtbl;
    pv.tbl11 = P->dtbl11;
    pv.ntracer = ntracer;
    pv.eta1_is_one = (mod_type1 == NGRTD_MOD_EXPONENTIAL);
    pv.eta2_is_one = (mod_type2 == NGRTD_MOD_EXPONENTIAL);
    pv.default_log10J = std::log10(j_flux(1., 2700, 1000, 3.0, 10.0, 0.05));   // run_age_mcmc_utils.py:90-91
    pv.tpl = (double)(L - 1) + dtp + ((L == 1) ? 1e-5 : 0.0);
    pv.itpl = 1.0 / pv.tpl;
    *out = P;
    return NGRTD_OK;
}

extern "C" int ngrtd_plan_destroy(ngrtd_plan* P) {
    if (!P) return NGRTD_OK;
    cudaFree(P->dXf); cudaFree(P->dXd); cudaFree(P->ditp); cudaFree(P->dxraw); cudaFree(P->dxrawd);
    cudaFree(P->dtbl); cudaFree(P->dtbl11);
    cudaFree(P->w_theta); cudaFree(P->w_out); cudaFree(P->w_logp); cudaFree(P->w_nu);
    if (P->hstream) cudaStreamDestroy(P->hstream);
    if (P->hstream2) cudaStreamDestroy(P->hstream2);
    if (P->hstream3) cudaStreamDestroy(P->hstream3);
    for (int i = 0; i < HOST_PARTS_MAX; i++) {
        if (P->ev_in[i]) cudaEventDestroy(P->ev_in[i]);
        if (P->ev_k[i]) cudaEventDestroy(P->ev_k[i]);
    }
    for (auto& sl : P->slots) {
        if (sl.busy && sl.ev_done) cudaEventSynchronize(sl.ev_done);
        cudaFree(sl.theta); cudaFree(sl.nu); cudaFree(sl.logp); cudaFree(sl.out);
        if (sl.ev_in) cudaEventDestroy(sl.ev_in);
        if (sl.ev_k) cudaEventDestroy(sl.ev_k);
        if (sl.ev_done) cudaEventDestroy(sl.ev_done);
        if (sl.s_k) cudaStreamDestroy(sl.s_k);
    }
    delete P;
    return NGRTD_OK;
}

extern "C" int ngrtd_plan_ntracer(const ngrtd_plan* P) { return P ? P->pv.ntracer : NGRTD_EINVAL; }

// ------------------------------------------------------------------------------------------- forward dispatch
#endif  // NGRTD_HAS_API

struct FwdTune { int warps, nt, ua; };

static FwdTune env_tune() {
    auto geti = [](const char* k) { const char* v = getenv(k); return v ? atoi(v) : 0; };
    return FwdTune{geti("NGRTD_FWD_WARPS"), geti("NGRTD_FWD_NT"), geti("NGRTD_FWD_UA")};
}

static int pick_warps(long long nunits, int nsm, int maxw) {
    // static schedule: what matters is warps per sub-partition; use the maximum the register budget allows
    // unless there are fewer units than warps.
    long long per_cta = (nunits + nsm - 1) / nsm;
    int w = maxw;
    while (w > 4 && (long long)(w - 4) >= per_cta) w -= 4;
    if (per_cta < 4) w = (int)std::max<long long>(1, per_cta);
    return w;
}

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-device (per-context) attribute of a kernel: remember what was
// configured for every device ordinal, so that a second plan on another GPU of the same thread configures its own context.
constexpr int MAX_DEVICES = 64;
struct SmemConfigured { size_t bytes[MAX_DEVICES] = {}; };
constexpr size_t SMEM_LIMIT = 227 * 1024;       // opt-in dynamic shared memory per CTA on sm_100

static bool pdl_enabled() {
    static const bool on = [] { const char* e = getenv("NGRTD_PDL"); return !(e && atoi(e) == 0); }();
    return on;
}

#ifndef NGRTD_FWD_MAXW
#define NGRTD_FWD_MAXW 16
#endif
constexpr int FWD_NT = 2, FWD_UA = 1, FWD_MAXW = NGRTD_FWD_MAXW;

// entry points of the kernel parts (defined in part 1 / part 2 of a partitioned build)
struct ngrtd_sampler;
int ngrtd_part_forward(ngrtd_plan* P, const SlotMap& sm, const double* theta, long long B, double* out, double* logp,
                       const LikPar& lik, cudaStream_t st, int stage);
int ngrtd_part_mcmc_age(ngrtd_sampler* S, const RunArgs& ra, cudaStream_t st);

#if NGRTD_HAS_FWD    // ---- part 1: k_forward instantiations
template <int C1, int C2, bool DYN, int NT, int UA, int MAXW, bool TAIL>
static int launch_forward_tt(ngrtd_plan* P, const SlotMap& sm, const double* theta, long long B, double* out,
                             double* logp, const LikPar& lik, cudaStream_t st, int stage, int warps_req) {
    using WT = WarpTiles<C1, C2, DYN, NT, UA, FWD_TB, TAIL ? 1 : 0, true>;
    const int Lloop = (TAIL && tail_active(P->pv, WT::ANY_G, WT::ANY_D)) ? P->pv.Kc : P->Lpad;
    long long nunits = (B + NT * 8 - 1) / (NT * 8);
    if (sm.ndim > NSLOT) stage = 0;                                  // staging slots are sized for <= NSLOT columns
    int warps = warps_req > 0 ? std::min(warps_req, MAXW) : MAXW;    // tape schedule: every warp takes an equal share
    if (warps > 4) warps &= ~3;
    auto total = [&](int w, int lc, bool tape) -> size_t {
        size_t sh = (size_t)fwd_smem_doubles<WT>(w, lc, tape);
        if (stage) sh = ((sh + 1) & ~(size_t)1) + (size_t)w * (NT * 8 * sm.ndim + 1) + 1;
        return sh * sizeof(double);
    };
    int lc_cap = WT::ANY_LOOP ? std::min(Lloop, LC_MAX) : 0;
    bool resident = !WT::ANY_LOOP || Lloop <= lc_cap;               // FwdCta::setup: tape <=> one chunk
    size_t sh = total(warps, lc_cap, resident);
    if (resident && sh > SMEM_LIMIT) {          // tables + hand-off slots of the tape do not fit: stream two chunks in lock step
        lc_cap = (Lloop / 2 + 3) & ~3;
        resident = false;
    }
    int grid;
    if (resident) {
        grid = (int)std::min<long long>(nunits, P->nsm);
        const long long per_cta = (nunits + grid - 1) / std::max(grid, 1);
        if (per_cta * std::max(1, Lloop / 4) >= (1LL << 31))
            return fail(NGRTD_EINVAL, "forward: batch too large for one launch (split it)");
    } else {
        if (warps_req <= 0) warps = pick_warps(nunits, P->nsm, MAXW);
        if (warps > 4) warps &= ~3;
        long long want = (nunits + warps - 1) / warps;
        grid = (int)std::min<long long>(want, P->nsm);
    }
    if (grid < 1) grid = 1;
    sh = total(warps, lc_cap, resident);
    if (sh > SMEM_LIMIT) return fail(NGRTD_EINVAL, "forward: shared-memory layout exceeds 227 KB");
    auto kern = k_forward<C1, C2, DYN, NT, UA, MAXW, TAIL>;
    static thread_local SmemConfigured configured;
    const int dev = (P->device >= 0 && P->device < MAX_DEVICES) ? P->device : 0;
    if (configured.bytes[dev] < sh) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        configured.bytes[dev] = sh;
    }
    // Programmatic dependent launch: launch i+1 may begin its set-up (table loads) while launch i drains; the kernel
    // waits for its predecessor (griddepcontrol.wait) before it reads theta or writes anything.
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)(warps * 32));
    cfg.dynamicSmemBytes = sh;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    static const int tape_min = [] { const char* e = getenv("NGRTD_TAPE_MIN"); return e ? atoi(e) : TAPE_MIN_GROUPS; }();
    CUDA_TRY(cudaLaunchKernelEx(&cfg, kern, P->pv, sm, theta, B, out, logp, lik, lc_cap, stage, tape_min));
    return NGRTD_OK;
}

// two instantiations per model pair: with and without the analytic-tail code (WarpTiles TM)
template <int C1, int C2, bool DYN, int NT, int UA, int MAXW>
static int launch_forward_t(ngrtd_plan* P, const SlotMap& sm, const double* theta, long long B, double* out,
                            double* logp, const LikPar& lik, cudaStream_t st, int stage, int warps_req) {
    constexpr bool ANY_G = (C1 == CLS_G || C2 == CLS_G), ANY_D = (C1 == CLS_D || C2 == CLS_D);
    if constexpr (ANY_G || ANY_D) {
        if (tail_active(P->pv, ANY_G, ANY_D))
            return launch_forward_tt<C1, C2, DYN, NT, UA, MAXW, true>(P, sm, theta, B, out, logp, lik, st, stage, warps_req);
    }
    return launch_forward_tt<C1, C2, DYN, NT, UA, MAXW, false>(P, sm, theta, B, out, logp, lik, st, stage, warps_req);
}


template <int C1, int C2, bool DYN>
static int launch_forward(ngrtd_plan* P, const SlotMap& sm, const double* theta, long long B, double* out,
                          double* logp, const LikPar& lik, cudaStream_t st, int stage) {
    FwdTune t = env_tune();
#ifdef NGRTD_TUNE
    // development build: every (NT, UA) variant of the looped kernels is compiled and selectable by env var
    if constexpr (!DYN && C1 != CLS_P && C2 != CLS_P) {
        int mw = (t.warps > 16) ? 24 : (t.warps > 12) ? 16 : (t.warps > 8) ? 12 : 8;
#define NGRTD_VARIANT(NT_, UA_, MW_) \
        if (t.nt == NT_ && t.ua == UA_ && mw == MW_) return launch_forward_t<C1, C2, DYN, NT_, UA_, MW_>(P, sm, theta, B, out, logp, lik, st, stage, t.warps);
        NGRTD_VARIANT(1, 2, 8) NGRTD_VARIANT(2, 1, 8) NGRTD_VARIANT(2, 2, 8) NGRTD_VARIANT(4, 1, 8)
        NGRTD_VARIANT(1, 1, 16) NGRTD_VARIANT(1, 2, 16) NGRTD_VARIANT(2, 2, 16) NGRTD_VARIANT(3, 1, 16)
        NGRTD_VARIANT(1, 1, 24) NGRTD_VARIANT(2, 1, 24) NGRTD_VARIANT(2, 1, 12) NGRTD_VARIANT(3, 1, 12) NGRTD_VARIANT(4, 1, 12)
#undef NGRTD_VARIANT
    }
#endif
    return launch_forward_t<C1, C2, DYN, FWD_NT, FWD_UA, FWD_MAXW>(P, sm, theta, B, out, logp, lik, st, stage, t.warps);
}

template <int C1, bool DYN>
static int dispatch_c2(ngrtd_plan* P, const SlotMap& sm, const double* theta, long long B, double* out, double* logp,
                       const LikPar& lik, cudaStream_t st, int stage) {
#ifdef NGRTD_EXP   /* timing-experiment build: only the three looped kernels of the benchmark */
    if (P->cls2 == CLS_NONE) return launch_forward<C1, CLS_NONE, DYN>(P, sm, theta, B, out, logp, lik, st, stage);
    if constexpr (C1 == CLS_G) if (P->cls2 == CLS_D) return launch_forward<C1, CLS_D, DYN>(P, sm, theta, B, out, logp, lik, st, stage);
    return fail(NGRTD_EINVAL, "experiment build: model pair not compiled");
#else
    switch (P->cls2) {
        case CLS_NONE: return launch_forward<C1, CLS_NONE, DYN>(P, sm, theta, B, out, logp, lik, st, stage);
        case CLS_P: return launch_forward<C1, CLS_P, DYN>(P, sm, theta, B, out, logp, lik, st, stage);
        case CLS_G: return launch_forward<C1, CLS_G, DYN>(P, sm, theta, B, out, logp, lik, st, stage);
        case CLS_D: return launch_forward<C1, CLS_D, DYN>(P, sm, theta, B, out, logp, lik, st, stage);
    }
    return fail(NGRTD_EINVAL, "bad model class");
#endif
}

template <bool DYN>
static int dispatch_c1(ngrtd_plan* P, const SlotMap& sm, const double* theta, long long B, double* out, double* logp,
                       const LikPar& lik, cudaStream_t st, int stage) {
    switch (P->cls1) {
#ifdef NGRTD_EXP
        case CLS_P: return fail(NGRTD_EINVAL, "experiment build: model pair not compiled");
#else
        case CLS_P: return dispatch_c2<CLS_P, DYN>(P, sm, theta, B, out, logp, lik, st, stage);
#endif
        case CLS_G: return dispatch_c2<CLS_G, DYN>(P, sm, theta, B, out, logp, lik, st, stage);
        case CLS_D: return dispatch_c2<CLS_D, DYN>(P, sm, theta, B, out, logp, lik, st, stage);
    }
    return fail(NGRTD_EINVAL, "bad model class");
}

int ngrtd_part_forward(ngrtd_plan* P, const SlotMap& sm, const double* theta, long long B, double* out, double* logp,
                       const LikPar& lik, cudaStream_t st, int stage) {
    // the per-chain-lambda path is only needed when thalf_cfc is actually sampled (run_age_mcmc_utils.py:107:
    // `'thalf_cfc' in self.p_names`); otherwise those tracers fall back to lambda = 0 through the same path.
#ifndef NGRTD_EXP
    if (P->dyn) return dispatch_c1<true>(P, sm, theta, B, out, logp, lik, st, stage);
#endif
    return dispatch_c1<false>(P, sm, theta, B, out, logp, lik, st, stage);
}
#endif  // NGRTD_HAS_FWD

#if NGRTD_HAS_API    // ---- part 0: C ABI (forward entry points, RTD weights, CE, likelihood)
static int make_slotmap(SlotMap& sm, int ndim, const int32_t* slot_of_col, bool dyn) {
    if (ndim < 1 || ndim > 32 || !slot_of_col) return fail(NGRTD_EINVAL, "theta: ndim must be in 1..32 with a slot map");
    sm.ndim = ndim;
    for (int s = 0; s < NSLOT; s++) sm.col_of_slot[s] = -1;
    sm.f2_complement = 0;
    for (int i = 0; i < ndim; i++) {
        int s = slot_of_col[i];
        if (s == NGRTD_P_F1_COMPLEMENT) {           // f1 column that also defines f2 = 1 - f1
            if (sm.col_of_slot[NGRTD_P_F1] >= 0 || sm.col_of_slot[NGRTD_P_F2] >= 0 || sm.f2_complement)
                return fail(NGRTD_EINVAL, "theta: the f1-complement column excludes f1 / f2 columns");
            sm.f2_complement = 1;
            s = NGRTD_P_F1;
        } else if (s == NGRTD_P_F1 || s == NGRTD_P_F2) {
            if (sm.f2_complement) return fail(NGRTD_EINVAL, "theta: the f1-complement column excludes f1 / f2 columns");
        }
        if (s < 0 || s >= NSLOT) return fail(NGRTD_EINVAL, "theta: unknown parameter slot " + std::to_string(s));
        sm.col_of_slot[s] = (signed char)i;
    }
    if (sm.col_of_slot[NGRTD_P_TAU1] < 0) return fail(NGRTD_EINVAL, "theta: tau1 is required");
    (void)dyn;
    return NGRTD_OK;
}

static int forward_common(ngrtd_plan* P, const double* theta, long long B, int ndim, const int32_t* slot_of_col,
                          double* out, double* logp, const LikPar& lik, cudaStream_t st, int stage) {
    if (!P) return fail(NGRTD_EINVAL, "null plan");
    if (B < 0) return fail(NGRTD_EINVAL, "B < 0");
    if (B == 0) return NGRTD_OK;
    if (!theta) return fail(NGRTD_EINVAL, "theta is null");
    DeviceGuard guard(P->device);          // launch on the plan's device whatever the caller's current device is
    if (!guard.ok()) return fail(NGRTD_ECUDA, "forward: cudaSetDevice(plan device) failed");
    SlotMap sm;
    int rc = make_slotmap(sm, ndim, slot_of_col, P->dyn);
    if (rc) return rc;
    return ngrtd_part_forward(P, sm, theta, B, out, logp, lik, st, stage);
}

// Parameter staging mode of k_forward: 0 = per-lane global loads, 1 = TMA bulk copy per unit with prefetch of the next
// unit, 2 = staged with plain lane loads (see k_forward).  NGRTD_STAGE overrides (development / A-B measurements).
static int default_stage() {
    const char* e = getenv("NGRTD_STAGE");
    return e ? atoi(e) : NGRTD_STAGE_DEFAULT;
}

static int forward_dev_stage(ngrtd_plan* P, const double* theta_d, int64_t B, int32_t ndim, const int32_t* slot_of_col,
                             double* out_d, void* stream, int stage) {
    if (!out_d) return fail(NGRTD_EINVAL, "out is null");
    LikPar lik{};
    lik.kind = -1;
    return forward_common(P, theta_d, B, ndim, slot_of_col, out_d, nullptr, lik, (cudaStream_t)stream, stage);
}

extern "C" int ngrtd_forward_dev(ngrtd_plan* P, const double* theta_d, int64_t B, int32_t ndim,
                                 const int32_t* slot_of_col, double* out_d, void* stream) {
    return forward_dev_stage(P, theta_d, B, ndim, slot_of_col, out_d, stream, default_stage());
}

static int fill_lik(LikPar& lik, const ngrtd_plan* P, int kind, const double* obs_mu, const double* obs_sd,
                    const double* nu_d) {
    if (kind != NGRTD_LIK_NORMAL && kind != NGRTD_LIK_STUDENTT) return fail(NGRTD_EINVAL, "unknown likelihood kind");
    if (!obs_mu || !obs_sd) return fail(NGRTD_EINVAL, "obs_mu / obs_sd is null");
    if (kind == NGRTD_LIK_STUDENTT && !nu_d) return fail(NGRTD_EINVAL, "student-t needs nu");
    lik.kind = kind;
    for (int t = 0; t < MAX_TRACER; t++) {
        double sd = t < P->pv.ntracer ? obs_sd[t] : 1.0;
        lik.obs[t] = t < P->pv.ntracer ? obs_mu[t] : 0.0;
        lik.isd[t] = 1.0 / sd;
        lik.lc[t] = kind == NGRTD_LIK_NORMAL ? -0.5 * std::log(2.0 * M_PI * sd * sd) : -std::log(sd);
    }
    lik.nu = nu_d;
    return NGRTD_OK;
}

static int forward_loglik_dev_stage(ngrtd_plan* P, const double* theta_d, int64_t B, int32_t ndim,
                                    const int32_t* slot_of_col, int32_t lik_kind, const double* obs_mu,
                                    const double* obs_sd, const double* nu_d, double* logp_d, double* model_out_d,
                                    void* stream, int stage) {
    if (!P) return fail(NGRTD_EINVAL, "null plan");
    if (!logp_d) return fail(NGRTD_EINVAL, "logp is null");
    LikPar lik{};
    int rc = fill_lik(lik, P, lik_kind, obs_mu, obs_sd, nu_d);
    if (rc) return rc;
    return forward_common(P, theta_d, B, ndim, slot_of_col, model_out_d, logp_d, lik, (cudaStream_t)stream, stage);
}

extern "C" int ngrtd_forward_loglik_dev(ngrtd_plan* P, const double* theta_d, int64_t B, int32_t ndim,
                                        const int32_t* slot_of_col, int32_t lik_kind, const double* obs_mu,
                                        const double* obs_sd, const double* nu_d, double* logp_d,
                                        double* model_out_d, void* stream) {
    return forward_loglik_dev_stage(P, theta_d, B, ndim, slot_of_col, lik_kind, obs_mu, obs_sd, nu_d, logp_d, model_out_d,
                                    stream, default_stage());
}

static int grow(double** p, size_t* have, size_t need) {
    if (*have >= need) return NGRTD_OK;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *have = 0;
    cudaError_t e = cudaMalloc((void**)p, need * sizeof(double));
    if (e != cudaSuccess) return fail(NGRTD_ENOMEM, std::string("workspace: ") + cudaGetErrorString(e));
    *have = need;
    return NGRTD_OK;
}

// Host-buffer path.  Large batches are split in two halves on two streams so that the host->device copy of the second
// half and the device->host copy of the first overlap with the kernels (the copies are ~40 % of a serial call at the
// cfg-3 batch: 3.7 MB in, 0.5 MB out over PCIe against a 0.12 ms kernel).  Pinned host memory is needed for the overlap;
// pageable buffers still work (the runtime stages them).
// device-visible alias of a pinned / registered host pointer (UVA), or nullptr for pageable memory
static const double* mapped_host_ptr(const double* h) {
    if (!h) return nullptr;
    cudaPointerAttributes a{};
    if (cudaPointerGetAttributes(&a, h) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
    return static_cast<const double*>(a.devicePointer);
}

static int forward_host_common(ngrtd_plan* P, const double* theta_h, int64_t B, int32_t ndim, const int32_t* slot_of_col,
                               int want_lik, int32_t lik_kind, const double* obs_mu, const double* obs_sd,
                               const double* nu_h, double* logp_h, double* model_out_h) {
    if (!P) return fail(NGRTD_EINVAL, "null plan");
    if (!theta_h) return fail(NGRTD_EINVAL, "null host buffer");
    if (B <= 0) return B == 0 ? NGRTD_OK : fail(NGRTD_EINVAL, "B < 0");
    DeviceGuard guard(P->device);
    if (!guard.ok()) return fail(NGRTD_ECUDA, "cudaSetDevice(plan device) failed");
    const int nt = P->pv.ntracer;
    int rc;
    const bool need_nu = want_lik && lik_kind == NGRTD_LIK_STUDENTT;
    if (need_nu && !nu_h) return fail(NGRTD_EINVAL, "student-t needs nu");
    // ---- zero-copy mode: pinned (page-locked, mapped) host buffers are handed to ONE kernel launch directly.  The
    // kernel pulls each unit's parameter rows over PCIe with a TMA bulk copy (prefetching the next unit under the
    // current one) and writes logp straight into the caller's buffer, so there is no staged copy in front of the first
    // unit or behind the last one.  model_out (scattered 8-byte stores) still goes through a device buffer.
    {
        const char* mode_s = getenv("NGRTD_HOST_MODE");
        const bool allow_mapped = !(mode_s && strcmp(mode_s, "copy") == 0);
        const double* theta_m = allow_mapped ? mapped_host_ptr(theta_h) : nullptr;
        const double* nu_m = (allow_mapped && need_nu) ? mapped_host_ptr(nu_h) : nullptr;
        if (theta_m && ndim <= NSLOT && (!need_nu || nu_m)) {
            double* logp_m = want_lik ? const_cast<double*>(mapped_host_ptr(logp_h)) : nullptr;
            if (want_lik && !logp_m && (rc = grow(&P->w_logp, &P->w_logp_n, (size_t)B))) return rc;
            if (model_out_h && (rc = grow(&P->w_out, &P->w_out_n, (size_t)B * nt))) return rc;
            cudaStream_t st = P->hstream2;
            const char* stg_s = getenv("NGRTD_STAGE_MAPPED");
            const int stg = stg_s ? atoi(stg_s) : 1;
            if (want_lik)
                rc = forward_loglik_dev_stage(P, theta_m, B, ndim, slot_of_col, lik_kind, obs_mu, obs_sd, nu_m,
                                              logp_m ? logp_m : P->w_logp, model_out_h ? P->w_out : nullptr, st, stg);
            else
                rc = forward_dev_stage(P, theta_m, B, ndim, slot_of_col, P->w_out, st, stg);
            if (rc) return rc;
            if (want_lik && !logp_m)
                CUDA_TRY(cudaMemcpyAsync(logp_h, P->w_logp, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, st));
            if (model_out_h)
                CUDA_TRY(cudaMemcpyAsync(model_out_h, P->w_out, (size_t)B * nt * sizeof(double), cudaMemcpyDeviceToHost, st));
            CUDA_TRY(cudaStreamSynchronize(st));
            return NGRTD_OK;
        }
    }
    if ((rc = grow(&P->w_theta, &P->w_theta_n, (size_t)B * ndim))) return rc;
    if (want_lik && (rc = grow(&P->w_logp, &P->w_logp_n, (size_t)B))) return rc;
    if (model_out_h && (rc = grow(&P->w_out, &P->w_out_n, (size_t)B * nt))) return rc;
    if (need_nu && (rc = grow(&P->w_nu, &P->w_nu_n, (size_t)B))) return rc;
    // Pipeline: all parts are copied in back to back on the copy-in stream; the compute stream runs part c as soon as
    // its event fires; the copy-out stream drains part c behind its kernel.  The copy engines and the SMs then work
    // concurrently for all but the first copy-in and the last copy-out.  Parts are unit-aligned (16 chains).  Measured
    // at the cfg-3 batch (65,536 chains): 1 part 0.226 ms, 2 parts 0.208, 4 parts 0.208, 6-8 parts 0.26-0.28 (a part
    // below ~16k chains no longer fills the persistent grid of 2,368 warps x 16 chains).
    const char* parts_s = getenv("NGRTD_HOST_PARTS");
    const int parts_env = parts_s ? atoi(parts_s) : 0;
    int nparts = parts_env > 0 ? parts_env : (int)std::min<int64_t>(2, B / 16384);
    nparts = std::max(1, std::min(nparts, HOST_PARTS_MAX));
    const int64_t part = nparts > 1 ? (((B + nparts - 1) / nparts + 15) & ~15LL) : B;
    cudaStream_t s_in = P->hstream, s_k = P->hstream2, s_out = P->hstream3;
    for (int c = 0; c < nparts; c++) {
        const int64_t b0 = c * part, n = std::min<int64_t>(part, B - b0);
        if (n <= 0) break;
        CUDA_TRY(cudaMemcpyAsync(P->w_theta + b0 * ndim, theta_h + b0 * ndim, (size_t)n * ndim * sizeof(double),
                                 cudaMemcpyHostToDevice, s_in));
        if (need_nu) CUDA_TRY(cudaMemcpyAsync(P->w_nu + b0, nu_h + b0, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, s_in));
        CUDA_TRY(cudaEventRecord(P->ev_in[c], s_in));
    }
    for (int c = 0; c < nparts; c++) {
        const int64_t b0 = c * part, n = std::min<int64_t>(part, B - b0);
        if (n <= 0) break;
        CUDA_TRY(cudaStreamWaitEvent(s_k, P->ev_in[c], 0));
        if (want_lik)
            rc = ngrtd_forward_loglik_dev(P, P->w_theta + b0 * ndim, n, ndim, slot_of_col, lik_kind, obs_mu, obs_sd,
                                          need_nu ? P->w_nu + b0 : nullptr, P->w_logp + b0,
                                          model_out_h ? P->w_out + b0 * nt : nullptr, s_k);
        else
            rc = ngrtd_forward_dev(P, P->w_theta + b0 * ndim, n, ndim, slot_of_col, P->w_out + b0 * nt, s_k);
        if (rc) return rc;
        CUDA_TRY(cudaEventRecord(P->ev_k[c], s_k));
        CUDA_TRY(cudaStreamWaitEvent(s_out, P->ev_k[c], 0));
        if (want_lik) CUDA_TRY(cudaMemcpyAsync(logp_h + b0, P->w_logp + b0, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, s_out));
        if (model_out_h)
            CUDA_TRY(cudaMemcpyAsync(model_out_h + b0 * nt, P->w_out + b0 * nt, (size_t)n * nt * sizeof(double),
                                     cudaMemcpyDeviceToHost, s_out));
    }
    CUDA_TRY(cudaStreamSynchronize(s_out));    // the last copy-out is behind every kernel and every copy-in
    return NGRTD_OK;
}

// ---- submit / wait: one batch per slot, copy-in / compute / copy-out on the plan's three streams (FIFO across slots), so
// the copies of neighbouring batches run under the kernel of the current one.
extern "C" int ngrtd_forward_loglik_host_submit(ngrtd_plan* P, const double* theta_h, int64_t B, int32_t ndim,
                                                const int32_t* slot_of_col, int32_t lik_kind, const double* obs_mu,
                                                const double* obs_sd, const double* nu_h, double* logp_h,
                                                double* model_out_h, int32_t slot) {
    if (!P) return fail(NGRTD_EINVAL, "null plan");
    if (slot < 0 || slot >= NGRTD_HOST_SLOTS) return fail(NGRTD_EINVAL, "submit: slot out of range");
    if (!theta_h || !logp_h) return fail(NGRTD_EINVAL, "null host buffer");
    if (B < 0) return fail(NGRTD_EINVAL, "B < 0");
    auto& sl = P->slots[slot];
    if (sl.busy) return fail(NGRTD_EINVAL, "submit: slot " + std::to_string(slot) + " is busy (call ngrtd_host_wait first)");
    if (B == 0) return NGRTD_OK;
    const bool need_nu = lik_kind == NGRTD_LIK_STUDENTT;
    if (need_nu && !nu_h) return fail(NGRTD_EINVAL, "student-t needs nu");
    DeviceGuard guard(P->device);
    if (!guard.ok()) return fail(NGRTD_ECUDA, "cudaSetDevice(plan device) failed");
    const int nt = P->pv.ntracer;
    int rc;
    if ((rc = grow(&sl.theta, &sl.theta_n, (size_t)B * ndim))) return rc;
    if ((rc = grow(&sl.logp, &sl.logp_n, (size_t)B))) return rc;
    if (model_out_h && (rc = grow(&sl.out, &sl.out_n, (size_t)B * nt))) return rc;
    if (need_nu && (rc = grow(&sl.nu, &sl.nu_n, (size_t)B))) return rc;
    cudaStream_t s_in = P->hstream, s_k = sl.s_k, s_out = P->hstream3;
    CUDA_TRY(cudaMemcpyAsync(sl.theta, theta_h, (size_t)B * ndim * sizeof(double), cudaMemcpyHostToDevice, s_in));
    if (need_nu) CUDA_TRY(cudaMemcpyAsync(sl.nu, nu_h, (size_t)B * sizeof(double), cudaMemcpyHostToDevice, s_in));
    CUDA_TRY(cudaEventRecord(sl.ev_in, s_in));
    CUDA_TRY(cudaStreamWaitEvent(s_k, sl.ev_in, 0));
    rc = ngrtd_forward_loglik_dev(P, sl.theta, B, ndim, slot_of_col, lik_kind, obs_mu, obs_sd, need_nu ? sl.nu : nullptr,
                                  sl.logp, model_out_h ? sl.out : nullptr, s_k);
    if (rc) {                                   // the copy-in is already enqueued on the caller's buffers: drain it
        cudaStreamSynchronize(s_in);
        return rc;
    }
    CUDA_TRY(cudaEventRecord(sl.ev_k, s_k));
    CUDA_TRY(cudaStreamWaitEvent(s_out, sl.ev_k, 0));
    CUDA_TRY(cudaMemcpyAsync(logp_h, sl.logp, (size_t)B * sizeof(double), cudaMemcpyDeviceToHost, s_out));
    if (model_out_h)
        CUDA_TRY(cudaMemcpyAsync(model_out_h, sl.out, (size_t)B * nt * sizeof(double), cudaMemcpyDeviceToHost, s_out));
    CUDA_TRY(cudaEventRecord(sl.ev_done, s_out));
    sl.busy = true;      // the slot's buffers are reused only after ngrtd_host_wait(slot)
    return NGRTD_OK;
}

extern "C" int ngrtd_host_wait(ngrtd_plan* P, int32_t slot) {
    if (!P) return fail(NGRTD_EINVAL, "null plan");
    if (slot < 0 || slot >= NGRTD_HOST_SLOTS) return fail(NGRTD_EINVAL, "wait: slot out of range");
    auto& sl = P->slots[slot];
    if (!sl.busy) return NGRTD_OK;
    DeviceGuard guard(P->device);
    sl.busy = false;
    CUDA_TRY(cudaEventSynchronize(sl.ev_done));
    return NGRTD_OK;
}

extern "C" int ngrtd_forward_host(ngrtd_plan* P, const double* theta_h, int64_t B, int32_t ndim,
                                  const int32_t* slot_of_col, double* out_h) {
    if (!out_h) return fail(NGRTD_EINVAL, "null host buffer");
    return forward_host_common(P, theta_h, B, ndim, slot_of_col, 0, 0, nullptr, nullptr, nullptr, nullptr, out_h);
}

extern "C" int ngrtd_forward_loglik_host(ngrtd_plan* P, const double* theta_h, int64_t B, int32_t ndim,
                                         const int32_t* slot_of_col, int32_t lik_kind, const double* obs_mu,
                                         const double* obs_sd, const double* nu_h, double* logp_h,
                                         double* model_out_h) {
    if (!logp_h) return fail(NGRTD_EINVAL, "null host buffer");
    return forward_host_common(P, theta_h, B, ndim, slot_of_col, 1, lik_kind, obs_mu, obs_sd, nu_h, logp_h, model_out_h);
}

// ------------------------------------------------------------------------------------------- class-API helpers
// gen_g_tp(): materialised, normalised weights -- one CTA per chain, direct per-lag formulae of the reference
// (utils/convolution_integral_utils.py:178-196,270).  API-parity path (plots, g_tau=...), not the hot loop.
__global__ void k_rtd_weights(int mod, int L, double dtp, const double* __restrict__ tau_, const double* __restrict__ eta_,
                              const double* __restrict__ D_, double* __restrict__ g) {
    const long long b = blockIdx.x;
    const double tau = tau_[b];
    const double eta = eta_ ? eta_[b] : 1.0;
    const double D = D_ ? D_[b] : 0.0;
    double* row = g + b * (long long)L;
    __shared__ double red[32];
    __shared__ int s_ix;
    if (mod == NGRTD_MOD_PISTON) {
        if (threadIdx.x == 0) {
            Comp<CLS_P> c;
            c.init(tau, 0.0, 0.0, dtp, L);
            s_ix = c.ix;
        }
        __syncthreads();
        for (int k = threadIdx.x; k < L; k += blockDim.x) row[k] = (k == s_ix) ? 1.0 : 0.0;
        return;
    }
    double part = 0.0;
    const double thr = __dmul_rn(tau, __dsub_rn(1.0, __ddiv_rn(1.0, eta)));
    for (int k = threadIdx.x; k < L; k += blockDim.x) {
        double tp = ((k == 0) ? 1e-5 : (double)k) + dtp;
        double w;
        if (mod == NGRTD_MOD_EXPONENTIAL) {
            w = (1.0 / tau) * exp(-tp / tau);
        } else if (mod == NGRTD_MOD_EXP_PIST_FLOW) {
            w = (tp >= thr) ? (eta / tau) * exp(-(eta * tp / tau) + eta - 1.0) : 0.0;
        } else {
            double x = tp / tau;
            double f1 = (1.0 / tau) / sqrt(4.0 * 3.14159265358979323846 * D * x);
            double om = 1.0 - x;
            double f2 = (1.0 / x) * exp(-1.0 * ((om * om) / (4.0 * D * x)));
            w = f1 * f2;
        }
        row[k] = w;
        part += w;
    }
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) red[0] = v;
    }
    __syncthreads();
    const double S = red[0];
    for (int k = threadIdx.x; k < L; k += blockDim.x) row[k] = row[k] / S;
}

extern "C" int ngrtd_rtd_weights_dev(int32_t mod_type, int32_t L, double dtp, const double* tau_d, const double* eta_d,
                                     const double* D_d, int64_t B, double* g_d, void* stream) {
    if (cls_of(mod_type) <= 0) return fail(NGRTD_EINVAL, "rtd_weights: unknown mod_type " + std::to_string(mod_type));
    if (L < 1 || !tau_d || !g_d) return fail(NGRTD_EINVAL, "rtd_weights: bad arguments");
    if (mod_type == NGRTD_MOD_EXP_PIST_FLOW && !eta_d) return fail(NGRTD_EINVAL, "rtd_weights: exp_pist_flow needs eta");
    if (mod_type == NGRTD_MOD_DISPERSION && !D_d) return fail(NGRTD_EINVAL, "rtd_weights: dispersion needs D");
    if (B <= 0) return B == 0 ? NGRTD_OK : fail(NGRTD_EINVAL, "B < 0");
    k_rtd_weights<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(mod_type, L, dtp, tau_d,
                                                                 mod_type == NGRTD_MOD_EXPONENTIAL ? nullptr : eta_d, D_d, g_d);
    CUDA_TRY(cudaGetLastError());
    return NGRTD_OK;
}

// ---- fracture / matrix-diffusion RTD ('frac_inf_diff'): frac_rtd_numba_disp (conv utils :36-63, the only
// numba-compiled code of the reference) + the post-processing of gen_g_tp (:238-270).  One CTA per (chain, lag)
// evaluates the 1000-point log-spaced inner quadrature: advective dispersion RTD x matrix-diffusion retention kernel,
// the retention kernel normalised by its own trapezoid integral (:59).  ~200 FP64 instructions per quadrature point.
constexpr int FDM_NQ = 1000;

__device__ __forceinline__ double block_sum(double v, double* red) {
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
    return t;
}

// np.interp(x, tp[:n], fp[:n]) on the lag grid tp[0] = 1e-5 + dtp, tp[k] = k + dtp (numpy's arithmetic: slope*(x-xp[j])+fp[j])
__device__ __forceinline__ double fdm_interp(double x, const double* __restrict__ fp, int n, double dtp) {
    auto xp = [&](int k) { return ((k == 0) ? 1e-5 : (double)k) + dtp; };
    if (x < xp(0)) return fp[0];
    if (!(x < xp(n - 1))) return fp[n - 1];
    int j = (int)floor(x - dtp);
    j = max(0, min(j, n - 2));
    while (j > 0 && x < xp(j)) j--;
    while (j < n - 2 && !(x < xp(j + 1))) j++;
    if (x == xp(j)) return fp[j];
    double slope = (fp[j + 1] - fp[j]) / (xp(j + 1) - xp(j));
    return __dadd_rn(__dmul_rn(slope, x - xp(j)), fp[j]);
}

// fext == nullptr: dispersion advective RTD (frac_rtd_numba_disp, :36-63); else the caller's advective RTD on the lag
// grid, linearly interpolated (frac_rtd_numba, :66-97)
__global__ void __launch_bounds__(128) k_fdm_lag(int L, double dtp, const double* __restrict__ tau_, const double* __restrict__ D_,
                                                 const double* __restrict__ bbar_, const double* __restrict__ phi_,
                                                 const double* __restrict__ fext, double* __restrict__ f) {
    __shared__ double tadv_s[FDM_NQ], tret_s[FDM_NQ], fret_s[FDM_NQ], fadv_s[FDM_NQ];
    __shared__ double red[4];
    const int i = blockIdx.x;
    const long long b = blockIdx.y;
    if (i == 0) {                                    // f_t_tran[0] = 0 (:256)
        if (threadIdx.x == 0) f[b * (long long)L] = 0.0;
        return;
    }
    const double tau = fext ? 1.0 : tau_[b], D = fext ? 1.0 : D_[b], bbar = bbar_[b], phi = phi_[b];
    const double D_o = (2.3e-9) * 60 * 60 * 24 * 365;                        // :244
    const double kappa = phi * sqrt((D_o * (phi * phi)) * 1);                // :247
    const double T = (double)i + dtp;
    const double hi = log10(T - 1.e-6);
    const double step = (hi - (-6.0)) / (double)(FDM_NQ - 1);                // np.logspace -> linspace step
    for (int m = threadIdx.x; m < FDM_NQ; m += blockDim.x) {
        double y = (m < FDM_NQ - 1) ? __dadd_rn(__dmul_rn((double)m, step), -6.0) : hi;
        double tadv = exp10(y);
        double x = tadv / tau;
        double om = 1. - x;
        double fadv;
        if (fext) fadv = fdm_interp(tadv, fext, i + 1, dtp);                                                                           // :88
        else fadv = ((1. / tau) / (sqrt(4. * 3.14159265358979323846 * D * x))) * (1. / x) * exp(-1. * ((om * om) / (4. * D * x)));     // :34
        double tret = T - tadv;
        double Beta = tadv / bbar;
        double fret = (kappa * Beta) / (2 * sqrt(3.14159265358979323846) * pow(tret, 1.5)) *
                      exp((-1 * (kappa * kappa) * (Beta * Beta)) / (4 * tret));                                                        // :58
        tadv_s[m] = tadv; tret_s[m] = tret; fret_s[m] = fret; fadv_s[m] = fadv;
    }
    __syncthreads();
    double part = 0.0;
    for (int m = threadIdx.x; m < FDM_NQ - 1; m += blockDim.x) part += (tret_s[m] - tret_s[m + 1]) * (fret_s[m + 1] + fret_s[m]);
    const double N = 0.5 * block_sum(part, red);                             // _trapz(f_ret[::-1], t_ret[::-1])  (:59)
    part = 0.0;
    for (int m = threadIdx.x; m < FDM_NQ - 1; m += blockDim.x) {
        double a0 = (fret_s[m] / N) * fadv_s[m], a1 = (fret_s[m + 1] / N) * fadv_s[m + 1];
        part += (tadv_s[m + 1] - tadv_s[m]) * (a1 + a0);
    }
    const double I = 0.5 * block_sum(part, red);                             // _trapz(f_i, tadv)  (:62)
    if (threadIdx.x == 0) f[b * (long long)L + i] = I;
}

// normalise by trapz over tp_ (tp_[0] = 0), mean travel time, then g / g.sum()  (:256-270)
__global__ void __launch_bounds__(256) k_fdm_post(int L, double dtp, double* __restrict__ g, double* __restrict__ fm_mu) {
    __shared__ double red[8];
    const long long b = blockIdx.x;
    double* row = g + b * (long long)L;
    auto tpv = [&](int k) { return k == 0 ? 0.0 : (double)k + dtp; };
    double p0 = 0.0, p1 = 0.0, ps = 0.0;
    for (int k = threadIdx.x; k < L - 1; k += blockDim.x) {
        double dx = tpv(k + 1) - tpv(k);
        p0 += dx * (row[k + 1] + row[k]);
    }
    const double nrm = 0.5 * block_sum(p0, red);
    p0 = 0.0;
    for (int k = threadIdx.x; k < L - 1; k += blockDim.x) {
        double dx = tpv(k + 1) - tpv(k);
        double f0 = row[k] / nrm, f1 = row[k + 1] / nrm;
        p0 += dx * (f1 + f0);
        p1 += dx * (f1 * tpv(k + 1) + f0 * tpv(k));
    }
    for (int k = threadIdx.x; k < L; k += blockDim.x) ps += row[k] / nrm;
    const double t0 = block_sum(p0, red), t1 = block_sum(p1, red), S = block_sum(ps, red);
    if (threadIdx.x == 0 && fm_mu) fm_mu[b] = t1 / t0;
    __syncthreads();
    for (int k = threadIdx.x; k < L; k += blockDim.x) row[k] = (row[k] / nrm) / S;
}

extern "C" int ngrtd_rtd_weights_fdm_dev(int32_t L, double dtp, const double* tau_d, const double* D_d, const double* bbar_d,
                                         const double* phi_d, int64_t B, double* g_d, double* fm_mu_d, void* stream) {
    if (L < 2 || !tau_d || !D_d || !bbar_d || !phi_d || !g_d) return fail(NGRTD_EINVAL, "rtd_weights_fdm: bad arguments");
    if (B <= 0) return B == 0 ? NGRTD_OK : fail(NGRTD_EINVAL, "B < 0");
    if (B > 65535) return fail(NGRTD_EINVAL, "rtd_weights_fdm: at most 65,535 parameter sets per call");
    cudaStream_t st = (cudaStream_t)stream;
    k_fdm_lag<<<dim3((unsigned)L, (unsigned)B), 128, 0, st>>>(L, dtp, tau_d, D_d, bbar_d, phi_d, nullptr, g_d);
    CUDA_TRY(cudaGetLastError());
    k_fdm_post<<<(unsigned)B, 256, 0, st>>>(L, dtp, g_d, fm_mu_d);
    CUDA_TRY(cudaGetLastError());
    return NGRTD_OK;
}

extern "C" int ngrtd_rtd_weights_fdm_ext_dev(int32_t L, double dtp, const double* f_tadv_ext_d, const double* bbar_d,
                                             const double* phi_d, int64_t B, double* g_d, double* fm_mu_d, void* stream) {
    if (L < 2 || !f_tadv_ext_d || !bbar_d || !phi_d || !g_d) return fail(NGRTD_EINVAL, "rtd_weights_fdm_ext: bad arguments");
    if (B <= 0) return B == 0 ? NGRTD_OK : fail(NGRTD_EINVAL, "B < 0");
    if (B > 65535) return fail(NGRTD_EINVAL, "rtd_weights_fdm_ext: at most 65,535 parameter sets per call");
    cudaStream_t st = (cudaStream_t)stream;
    k_fdm_lag<<<dim3((unsigned)L, (unsigned)B), 128, 0, st>>>(L, dtp, nullptr, nullptr, bbar_d, phi_d, f_tadv_ext_d, g_d);
    CUDA_TRY(cudaGetLastError());
    k_fdm_post<<<(unsigned)B, 256, 0, st>>>(L, dtp, g_d, fm_mu_d);
    CUDA_TRY(cudaGetLastError());
    return NGRTD_OK;
}

// convolve(g_tau=g) tail: decay/ingrowth (:313-316), input assembly (:320-333), dot (:336-337); one CTA per row
__global__ void k_convolve_g(int L, double dtp, const double* __restrict__ g, const double* __restrict__ series,
                             const double* __restrict__ lag_index, const double* __restrict__ lambda, int rad_accum,
                             const double* __restrict__ J, double* __restrict__ out) {
    const long long b = blockIdx.x;
    const double* row = g + b * (long long)L;
    const double lam = lambda ? lambda[b] : 0.0;
    const double Jb = (J && rad_accum == NGRTD_ACC_4HE) ? J[b] : 0.0;
    __shared__ double red[32];
    double part = 0.0;
    for (int k = threadIdx.x; k < L; k += blockDim.x) {
        double tp = ((k == 0) ? 1e-5 : (double)k) + dtp;
        double dec = exp(-lam * tp);
        double gd = row[k] * (rad_accum == NGRTD_ACC_3HE ? (1 - dec) : dec);
        double c = series[k];
        if (rad_accum == NGRTD_ACC_4HE) c = c + (lag_index ? lag_index[k] : (double)k) * Jb;
        part += c * gd;
    }
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = part;
    __syncthreads();
    if (threadIdx.x < 32) {
        double v = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (threadIdx.x == 0) out[b] = v;
    }
}

extern "C" int ngrtd_convolve_g_dev(int32_t L, double dtp, const double* g_d, int64_t B, const double* series_d,
                                    const double* lag_index_d, const double* lambda_d, int32_t rad_accum,
                                    const double* J_d, double* out_d, void* stream) {
    if (L < 1 || !g_d || !series_d || !out_d) return fail(NGRTD_EINVAL, "convolve_g: bad arguments");
    if (rad_accum < 0 || rad_accum > 2) return fail(NGRTD_EINVAL, "convolve_g: unknown rad_accum");
    if (B <= 0) return B == 0 ? NGRTD_OK : fail(NGRTD_EINVAL, "B < 0");
    k_convolve_g<<<(unsigned)B, 256, 0, (cudaStream_t)stream>>>(L, dtp, g_d, series_d, lag_index_d, lambda_d, rad_accum,
                                                                J_d, out_d);
    CUDA_TRY(cudaGetLastError());
    return NGRTD_OK;
}

// ------------------------------------------------------------------------------------------- CE model
static int make_gases(GasList& gl, int ngas, const int32_t* gases) {
    if (ngas < 1 || ngas > 5 || !gases) return fail(NGRTD_EINVAL, "ce: ngas must be in 1..5");
    gl.n = ngas;
    for (int i = 0; i < ngas; i++) {
        if (gases[i] < 0 || gases[i] > 4) return fail(NGRTD_EINVAL, "ce: gas id must be 0..4 (He,Ne,Ar,Kr,Xe)");
        gl.id[i] = gases[i];
    }
    return NGRTD_OK;
}

extern "C" int ngrtd_ce_dev(int32_t what, int32_t ngas, const int32_t* gases, const double* E_d, const double* T_d,
                            const double* Ae_d, const double* F_d, const double* P_d, double S, int64_t B,
                            double* out_d, void* stream) {
    GasList gl;
    int rc = make_gases(gl, ngas, gases);
    if (rc) return rc;
    if (what < 0 || what > 6) return fail(NGRTD_EINVAL, "ce: unknown output selector");
    if (!T_d || !out_d) return fail(NGRTD_EINVAL, "ce: T / out is null");
    if (!P_d && !E_d && what != 4 && what != 6) return fail(NGRTD_EINVAL, "ce: need E (lapse rate) or P");
    if (what <= 1 && (!Ae_d || !F_d)) return fail(NGRTD_EINVAL, "ce: ce_exc needs Ae and F");
    if (B <= 0) return B == 0 ? NGRTD_OK : fail(NGRTD_EINVAL, "B < 0");
    unsigned grid = (unsigned)((B + 127) / 128);
    k_ce<<<grid, 128, 0, (cudaStream_t)stream>>>(what, gl, E_d, T_d, Ae_d, F_d, P_d, S, B, out_d);
    CUDA_TRY(cudaGetLastError());
    return NGRTD_OK;
}

extern "C" int ngrtd_ce_wrapper_dev(int32_t ngas, const int32_t* gases, const double* theta_d, int64_t B,
                                    double* out_d, void* stream) {
    GasList gl;
    int rc = make_gases(gl, ngas, gases);
    if (rc) return rc;
    if (!theta_d || !out_d) return fail(NGRTD_EINVAL, "ce_wrapper: null pointer");
    if (B <= 0) return B == 0 ? NGRTD_OK : fail(NGRTD_EINVAL, "B < 0");
    unsigned grid = (unsigned)((B + 127) / 128);
    k_ce_wrapper<<<grid, 128, 0, (cudaStream_t)stream>>>(gl, theta_d, B, out_d);
    CUDA_TRY(cudaGetLastError());
    return NGRTD_OK;
}

extern "C" int ngrtd_ce_host(int32_t what, int32_t ngas, const int32_t* gases, const double* E_h, const double* T_h,
                             const double* Ae_h, const double* F_h, const double* P_h, double S, int64_t B,
                             double* out_h) {
    if (B <= 0) return B == 0 ? NGRTD_OK : fail(NGRTD_EINVAL, "B < 0");
    if (!T_h || !out_h) return fail(NGRTD_EINVAL, "ce: T / out is null");
    if (ngas < 1 || ngas > 5) return fail(NGRTD_EINVAL, "ce: ngas must be in 1..5");
    double* buf = nullptr;
    size_t n = (size_t)B;
    CUDA_TRY(cudaMalloc((void**)&buf, (5 + ngas) * n * sizeof(double)));
    const double* src[5] = {E_h, T_h, Ae_h, F_h, P_h};
    double* dev[5];
    for (int i = 0; i < 5; i++) {
        dev[i] = src[i] ? buf + i * n : nullptr;
        if (src[i]) {
            cudaError_t e = cudaMemcpy(dev[i], src[i], n * sizeof(double), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { cudaFree(buf); return fail(NGRTD_ECUDA, cudaGetErrorString(e)); }
        }
    }
    double* dout = buf + 5 * n;
    int rc = ngrtd_ce_dev(what, ngas, gases, dev[0], dev[1], dev[2], dev[3], dev[4], S, B, dout, nullptr);
    if (rc == NGRTD_OK) {
        cudaError_t e = cudaMemcpy(out_h, dout, (size_t)ngas * n * sizeof(double), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(NGRTD_ECUDA, cudaGetErrorString(e));
    }
    cudaFree(buf);
    return rc;
}

// ------------------------------------------------------------------------------------------- CFC / SF6 corrections
extern "C" int ngrtd_cfc_dev(int32_t what, int32_t nspecies, const int32_t* species, const double* E_d, const double* T_d,
                             const double* Ae_d, const double* F_d, const double* X_d, double S, int64_t B, double* out_d,
                             void* stream) {
    if (nspecies < 1 || nspecies > 4 || !species) return fail(NGRTD_EINVAL, "cfc: nspecies must be in 1..4");
    SpeciesList sl;
    sl.n = nspecies;
    for (int i = 0; i < nspecies; i++) {
        if (species[i] != 11 && species[i] != 12 && species[i] != 113 && species[i] != 6)
            return fail(NGRTD_EINVAL, "cfc: species must be 11, 12, 113 (CFCs) or 6 (SF6)");
        sl.id[i] = species[i];
    }
    if (what < 0 || what > 3) return fail(NGRTD_EINVAL, "cfc: unknown output selector");
    if (!T_d || !out_d) return fail(NGRTD_EINVAL, "cfc: T / out is null");
    if (what != 3 && (!E_d || !X_d)) return fail(NGRTD_EINVAL, "cfc: E and the concentration / mixing-ratio input are required");
    if ((what == 0 || what == 2) && (!Ae_d || !F_d)) return fail(NGRTD_EINVAL, "cfc: excess-air corrections need Ae and F");
    if (B <= 0) return B == 0 ? NGRTD_OK : fail(NGRTD_EINVAL, "B < 0");
    unsigned grid = (unsigned)((B + 127) / 128);
    k_cfc<<<grid, 128, 0, (cudaStream_t)stream>>>(what, sl, E_d, T_d, Ae_d, F_d, X_d, S, B, out_d);
    CUDA_TRY(cudaGetLastError());
    return NGRTD_OK;
}

extern "C" int ngrtd_cfc_host(int32_t what, int32_t nspecies, const int32_t* species, const double* E_h, const double* T_h,
                              const double* Ae_h, const double* F_h, const double* X_h, double S, int64_t B, double* out_h) {
    if (B <= 0) return B == 0 ? NGRTD_OK : fail(NGRTD_EINVAL, "B < 0");
    if (!T_h || !out_h) return fail(NGRTD_EINVAL, "cfc: T / out is null");
    if (nspecies < 1 || nspecies > 4) return fail(NGRTD_EINVAL, "cfc: nspecies must be in 1..4");
    const size_t n = (size_t)B, ns = (size_t)nspecies;
    double* buf = nullptr;
    CUDA_TRY(cudaMalloc((void**)&buf, (4 + 2 * ns) * n * sizeof(double)));
    const double* src[4] = {E_h, T_h, Ae_h, F_h};
    double* dev[4];
    for (int i = 0; i < 4; i++) {
        dev[i] = src[i] ? buf + i * n : nullptr;
        if (src[i]) {
            cudaError_t e = cudaMemcpy(dev[i], src[i], n * sizeof(double), cudaMemcpyHostToDevice);
            if (e != cudaSuccess) { cudaFree(buf); return fail(NGRTD_ECUDA, cudaGetErrorString(e)); }
        }
    }
    double* dX = X_h ? buf + 4 * n : nullptr;
    if (X_h) {
        cudaError_t e = cudaMemcpy(dX, X_h, ns * n * sizeof(double), cudaMemcpyHostToDevice);
        if (e != cudaSuccess) { cudaFree(buf); return fail(NGRTD_ECUDA, cudaGetErrorString(e)); }
    }
    double* dout = buf + (4 + ns) * n;
    int rc = ngrtd_cfc_dev(what, nspecies, species, dev[0], dev[1], dev[2], dev[3], dX, S, B, dout, nullptr);
    if (rc == NGRTD_OK) {
        cudaError_t e = cudaMemcpy(out_h, dout, ns * n * sizeof(double), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) rc = fail(NGRTD_ECUDA, cudaGetErrorString(e));
    }
    cudaFree(buf);
    return rc;
}

// ------------------------------------------------------------------------------------------- stand-alone loglik
struct ObsPar { int T; double obs[16]; double isd[16]; double lc[16]; };

__global__ void k_loglik(int kind, ObsPar op, const double* __restrict__ mu, const double* __restrict__ nu,
                         long long B, double* __restrict__ logp) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= B) return;
    double acc = 0.0;
    if (kind == NGRTD_LIK_STUDENTT) {
        double n = nu[i], cst = lik_studentt_const(n);
        for (int t = 0; t < op.T; t++) acc += lik_term_studentt(op.obs[t], mu[i * op.T + t], op.isd[t], op.lc[t], n, cst);
    } else {
        for (int t = 0; t < op.T; t++) acc += lik_term_normal(op.obs[t], mu[i * op.T + t], op.isd[t], op.lc[t]);
    }
    logp[i] = acc;
}

extern "C" int ngrtd_loglik_dev(int32_t lik_kind, int32_t T, const double* mu_d, const double* obs_mu,
                                const double* obs_sd, const double* nu_d, int64_t B, double* logp_d, void* stream) {
    if (lik_kind != NGRTD_LIK_NORMAL && lik_kind != NGRTD_LIK_STUDENTT) return fail(NGRTD_EINVAL, "unknown likelihood kind");
    if (T < 1 || T > 16) return fail(NGRTD_EINVAL, "loglik: T must be in 1..16");
    if (!mu_d || !obs_mu || !obs_sd || !logp_d) return fail(NGRTD_EINVAL, "loglik: null pointer");
    if (lik_kind == NGRTD_LIK_STUDENTT && !nu_d) return fail(NGRTD_EINVAL, "student-t needs nu");
    if (B <= 0) return B == 0 ? NGRTD_OK : fail(NGRTD_EINVAL, "B < 0");
    ObsPar op;
    op.T = T;
    for (int t = 0; t < 16; t++) {
        double sd = t < T ? obs_sd[t] : 1.0;
        op.obs[t] = t < T ? obs_mu[t] : 0.0;
        op.isd[t] = 1.0 / sd;
        op.lc[t] = lik_kind == NGRTD_LIK_NORMAL ? -0.5 * std::log(2.0 * M_PI * sd * sd) : -std::log(sd);
    }
    unsigned grid = (unsigned)((B + 127) / 128);
    k_loglik<<<grid, 128, 0, (cudaStream_t)stream>>>(lik_kind, op, mu_d, nu_d, B, logp_d);
    CUDA_TRY(cudaGetLastError());
    return NGRTD_OK;
}

// ------------------------------------------------------------------------------------------- sampler
#endif  // NGRTD_HAS_API

struct ngrtd_sampler {
    int device = 0;
    ngrtd_plan* plan = nullptr;
    SamplerView sv{};
    double tune_drop_fraction = 0.9;
    long long step = 0, ndraws = 0, hist_start = 0;
    size_t n_q = 0;
    double* d_groups = nullptr;     // [3, G, ntr]: obs, 1/sd, likelihood constant
    double* d_pool = nullptr;       // workspace of ngrtd_sampler_pooled_moments
};

#if NGRTD_HAS_API
static double lbeta(double a, double b) { return std::lgamma(a) + std::lgamma(b) - std::lgamma(a + b); }
#endif

// shared memory of k_mcmc_age<..., TB> for a lag chunk of lc_cap lags: forward tables + one record per resident chain + priors
#if NGRTD_HAS_MCMC   // ---- part 2: k_mcmc_age instantiations
template <class WT>
static size_t mcmc_age_smem(int warps, int lc_cap, int ndr) {
    size_t sh = (size_t)fwd_smem_doubles<WT>(warps, lc_cap, false);
    sh += (size_t)warps * WT::NTILES * 8 * ch_rec_doubles(ndr) + (sizeof(PriorDev) * ND_MAX + 7) / 8;
    return sh * sizeof(double);
}

template <int C1, int C2, bool DYN, bool TAIL, int TB, int NDR>
static int launch_mcmc_age_tb(ngrtd_sampler* S, const RunArgs& ra, cudaStream_t st, int warps, int lc_cap, size_t sh) {
    constexpr int NT = FWD_NT, UA = FWD_UA, MAXW = FWD_MAXW;
    ngrtd_plan* P = S->plan;
    const long long nunits = (S->sv.B + NT * 8 - 1) / (NT * 8);
    auto kern = k_mcmc_age<C1, C2, DYN, NT, UA, MAXW, TAIL, TB, NDR>;
    static thread_local SmemConfigured configured;
    const int dev = (S->device >= 0 && S->device < MAX_DEVICES) ? S->device : 0;
    if (configured.bytes[dev] < sh) {
        CUDA_TRY(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh));
        configured.bytes[dev] = sh;
    }
    long long want = (nunits + warps - 1) / warps;
    int grid = (int)std::max<long long>(1, std::min<long long>(want, P->nsm));
    kern<<<grid, warps * 32, sh, st>>>(P->pv, S->sv, ra, lc_cap);
    CUDA_TRY(cudaGetLastError());
    return NGRTD_OK;
}

static bool mcmc_big_table_enabled() {
    static const bool on = [] { const char* e = getenv("NGRTD_MCMC_TB11"); return !(e && atoi(e) == 0); }();
    return on;
}

template <int C1, int C2, bool DYN, bool TAIL>
static int launch_mcmc_age_t(ngrtd_sampler* S, const RunArgs& ra, cudaStream_t st) {
    constexpr int NT = FWD_NT, UA = FWD_UA, MAXW = FWD_MAXW;
    using WT = WarpTiles<C1, C2, DYN, NT, UA, MCMC_TB, TAIL ? 1 : 0>;
    ngrtd_plan* P = S->plan;
    const long long B = S->sv.B;
    long long nunits = (B + NT * 8 - 1) / (NT * 8);
    int warps = pick_warps(nunits, P->nsm, MAXW);
    if (warps > 4) warps &= ~3;
    const int Lloop = (TAIL && tail_active(P->pv, WT::ANY_G, WT::ANY_D)) ? P->pv.Kc : P->Lpad;
    int lc_cap = WT::ANY_LOOP ? std::min(Lloop, LC_MAX) : 0;
    // dispersion plans with up to ND_SMALL sampler dimensions: compact per-chain records and the 2,048-entry exp table
    // (quadratic, one DFMA less per weight) when that fits next to the lag tables without shrinking the resident chunk;
    // otherwise ND_MAX records and the 128-entry table (cubic).  Plans with a constant tail keep the small table: their lag
    // loop is a few dozen groups and the time is in the tail quadrature, so the second instantiation would only cost build time.
    if constexpr (WT::ANY_D && !TAIL) {
        using WTB = WarpTiles<C1, C2, DYN, NT, UA, MCMC_TB_BIG, TAIL ? 1 : 0>;
        const size_t shb = mcmc_age_smem<WTB>(warps, lc_cap, ND_SMALL);
        if (S->sv.nd <= ND_SMALL && shb <= SMEM_LIMIT && mcmc_big_table_enabled())
            return launch_mcmc_age_tb<C1, C2, DYN, TAIL, MCMC_TB_BIG, ND_SMALL>(S, ra, st, warps, lc_cap, shb);
    }
    // shrink the lag chunk until the layout fits
    size_t sh = 0;
    for (;;) {
        sh = mcmc_age_smem<WT>(warps, lc_cap, ND_MAX);
        if (sh <= SMEM_LIMIT || lc_cap <= 64) break;
        lc_cap = (lc_cap / 2 + 3) & ~3;
    }
    if (sh > SMEM_LIMIT) return fail(NGRTD_EINVAL, "sampler: shared-memory budget exceeded");
    return launch_mcmc_age_tb<C1, C2, DYN, TAIL, MCMC_TB, ND_MAX>(S, ra, st, warps, lc_cap, sh);
}

// two instantiations per model pair, as for k_forward: with and without the constant-tail code
template <int C1, int C2, bool DYN>
static int launch_mcmc_age(ngrtd_sampler* S, const RunArgs& ra, cudaStream_t st) {
    constexpr bool ANY_G = (C1 == CLS_G || C2 == CLS_G), ANY_D = (C1 == CLS_D || C2 == CLS_D);
    if constexpr (ANY_G || ANY_D) {
        if (tail_active(S->plan->pv, ANY_G, ANY_D)) return launch_mcmc_age_t<C1, C2, DYN, true>(S, ra, st);
    }
    return launch_mcmc_age_t<C1, C2, DYN, false>(S, ra, st);
}

template <int C1, bool DYN>
static int mcmc_c2(ngrtd_sampler* S, const RunArgs& ra, cudaStream_t st) {
#ifdef NGRTD_EXP
    if (S->plan->cls2 == CLS_NONE) return launch_mcmc_age<C1, CLS_NONE, DYN>(S, ra, st);
    if constexpr (C1 == CLS_G) if (S->plan->cls2 == CLS_D) return launch_mcmc_age<C1, CLS_D, DYN>(S, ra, st);
    return fail(NGRTD_EINVAL, "experiment build: model pair not compiled");
#else
    switch (S->plan->cls2) {
        case CLS_NONE: return launch_mcmc_age<C1, CLS_NONE, DYN>(S, ra, st);
        case CLS_P: return launch_mcmc_age<C1, CLS_P, DYN>(S, ra, st);
        case CLS_G: return launch_mcmc_age<C1, CLS_G, DYN>(S, ra, st);
        case CLS_D: return launch_mcmc_age<C1, CLS_D, DYN>(S, ra, st);
    }
    return fail(NGRTD_EINVAL, "bad model class");
#endif
}

template <bool DYN>
static int mcmc_c1(ngrtd_sampler* S, const RunArgs& ra, cudaStream_t st) {
    switch (S->plan->cls1) {
#ifndef NGRTD_EXP
        case CLS_P: return mcmc_c2<CLS_P, DYN>(S, ra, st);
#endif
        case CLS_G: return mcmc_c2<CLS_G, DYN>(S, ra, st);
        case CLS_D: return mcmc_c2<CLS_D, DYN>(S, ra, st);
    }
    return fail(NGRTD_EINVAL, "bad model class");
}

int ngrtd_part_mcmc_age(ngrtd_sampler* S, const RunArgs& ra, cudaStream_t st) {
#ifndef NGRTD_EXP
    if (S->plan->dyn) return mcmc_c1<true>(S, ra, st);
#endif
    return mcmc_c1<false>(S, ra, st);
}
#endif  // NGRTD_HAS_MCMC

#if NGRTD_HAS_API    // ---- part 0: C ABI (samplers, probes)
static int sampler_launch(ngrtd_sampler* S, const RunArgs& ra, cudaStream_t st) {
    if (S->sv.model == 1) {
        unsigned grid = (unsigned)((S->sv.B + 63) / 64);
        // register-resident instantiations for the dimension counts of the reference's model (nu_ sampled / fixed);
        // NGRTD_NG_GENERIC=1 forces the run-time-nd kernel (same trajectories, bit for bit)
        static const bool generic = [] { const char* e = getenv("NGRTD_NG_GENERIC"); return e && e[0] == '1'; }();
        if (!generic && (S->sv.nd == 6 || S->sv.nd == 5)) {
            // (static shared memory 7 ND x 64 doubles per block; the default carve-out is the fastest: asking for the
            //  maximum shrinks L1 and costs 40 %, measured)
            if (S->sv.nd == 6) k_mcmc_ng_r<6><<<grid, 64, 0, st>>>(S->sv, ra);
            else k_mcmc_ng_r<5><<<grid, 64, 0, st>>>(S->sv, ra);
        } else {
            k_mcmc_ng<<<grid, 64, 0, st>>>(S->sv, ra);
        }
        CUDA_TRY(cudaGetLastError());
        return NGRTD_OK;
    }
    return ngrtd_part_mcmc_age(S, ra, st);
}

extern "C" int ngrtd_sampler_destroy(ngrtd_sampler* S) {
    if (!S) return NGRTD_OK;
    SamplerView& v = S->sv;
    cudaFree(v.q); cudaFree(v.logp); cudaFree(v.lamb); cudaFree(v.scal); cudaFree(v.acc_win); cudaFree(v.acc_tot);
    cudaFree(v.hist); cudaFree(v.wf_mean); cudaFree(v.wf_m2); cudaFree(S->d_groups); cudaFree(S->d_pool);
    delete S;
    return NGRTD_OK;
}

extern "C" int ngrtd_sampler_create(ngrtd_sampler** out, const ngrtd_sampler_cfg* cfg, ngrtd_plan* plan, int64_t nchains,
                                    const double* q0, int32_t device) {
    if (!out || !cfg) return fail(NGRTD_EINVAL, "sampler: null pointer");
    *out = nullptr;
    if (cfg->ndim < 1 || cfg->ndim > ND_MAX) return fail(NGRTD_EINVAL, "sampler: ndim must be in 1..10");
    if (nchains < 1) return fail(NGRTD_EINVAL, "sampler: nchains must be >= 1");
    if (cfg->tune_interval < 1) return fail(NGRTD_EINVAL, "sampler: tune_interval must be >= 1");
    if (cfg->hist_cap < 2) return fail(NGRTD_EINVAL, "sampler: hist_cap must be >= 2");
    if (cfg->lik_kind != NGRTD_LIK_NORMAL && cfg->lik_kind != NGRTD_LIK_STUDENTT) return fail(NGRTD_EINVAL, "sampler: unknown likelihood");
    if (cfg->nobs < 1 || cfg->nobs > MAX_TRACER) return fail(NGRTD_EINVAL, "sampler: nobs must be in 1..8");
    if (plan && cfg->nobs != plan->pv.ntracer) return fail(NGRTD_EINVAL, "sampler: nobs must equal the plan's tracer count");
    if (!plan && (cfg->ngas < 1 || cfg->ngas > 5 || cfg->ngas != cfg->nobs)) return fail(NGRTD_EINVAL, "sampler: noble-gas model needs 1..5 gases = nobs");
    int sdev = plan ? plan->device : device;
    if (sdev < 0) CUDA_TRY(cudaGetDevice(&sdev));
    DeviceGuard guard(sdev);                // the caller's current device is restored on return
    if (!guard.ok()) return fail(NGRTD_ECUDA, "sampler: cudaSetDevice(" + std::to_string(sdev) + ") failed");
    auto* S = new ngrtd_sampler();
    S->device = sdev;
    S->plan = plan;
    S->tune_drop_fraction = cfg->tune_drop_fraction;
    SamplerView& v = S->sv;
    v.nd = cfg->ndim;
    v.model = plan ? 0 : 1;
    v.sampled_mask = 0;
    std::vector<double> start(cfg->ndim);
    for (int d = 0; d < cfg->ndim; d++) {
        const ngrtd_prior& pr = cfg->prior[d];
        PriorDev pd{pr.kind, pr.target, pr.p0, pr.p1, pr.lo, pr.hi, 0.0};
        if (pr.target < 0 || pr.target >= NVAL) { delete S; return fail(NGRTD_EINVAL, "sampler: prior target out of range"); }
        switch (pr.kind) {
            case NGRTD_PRIOR_UNIFORM:
                if (!(pr.p1 > pr.p0)) { delete S; return fail(NGRTD_EINVAL, "sampler: uniform needs p1 > p0"); }
                start[d] = 0.0;                                               // logit(1/2)
                break;
            case NGRTD_PRIOR_BETA: {
                if (!(pr.p0 > 0 && pr.p1 > 0)) { delete S; return fail(NGRTD_EINVAL, "sampler: beta needs positive shapes"); }
                pd.c = lbeta(pr.p0, pr.p1);
                double m = pr.p0 / (pr.p0 + pr.p1);                           // test value = mean
                start[d] = std::log(m / (1.0 - m));
                break;
            }
            case NGRTD_PRIOR_NORMAL:
                if (!(pr.p1 > 0)) { delete S; return fail(NGRTD_EINVAL, "sampler: normal needs sigma > 0"); }
                pd.c = -0.5 * std::log(2.0 * M_PI * pr.p1 * pr.p1);
                start[d] = pr.p0;
                break;
            case NGRTD_PRIOR_HALFNORMAL:
                if (!(pr.p0 > 0)) { delete S; return fail(NGRTD_EINVAL, "sampler: halfnormal needs sigma > 0"); }
                pd.c = 0.5 * std::log(2.0 / M_PI) - std::log(pr.p0);
                start[d] = std::log(pr.p0 * std::sqrt(2.0 / M_PI));           // test value = mean
                break;
            default:
                delete S;
                return fail(NGRTD_EINVAL, "sampler: unknown prior kind");
        }
        v.pr[d] = pd;
        v.sampled_mask |= 1u << pr.target;
        if (q0) start[d] = q0[d];
    }
    v.lik_kind = cfg->lik_kind;
    v.nu_sampled = cfg->nu_sampled;
    v.nu_lo = cfg->nu_lo; v.nu_hi = cfg->nu_hi; v.nu_fixed = cfg->nu_fixed;
    v.ntr = cfg->nobs;
    for (int t = 0; t < MAX_TRACER; t++) {
        double sd = t < cfg->nobs ? cfg->obs_sd[t] : 1.0;
        v.obs[t] = t < cfg->nobs ? cfg->obs_mu[t] : 0.0;
        v.isd[t] = 1.0 / sd;
        v.lc[t] = cfg->lik_kind == NGRTD_LIK_NORMAL ? -0.5 * std::log(2.0 * M_PI * sd * sd) : -std::log(sd);
    }
    v.f2_from_f1 = cfg->f2_from_f1;
    v.proposal_dist = cfg->proposal_dist;
    v.de_mcz = cfg->de_mcz;
    v.tune_target = cfg->tune_target;
    v.tune_interval = cfg->tune_interval;
    v.seed = cfg->seed;
    v.chain_offset = cfg->chain_offset;
    v.B = nchains;
    v.hist_cap = cfg->hist_cap;
    v.g_obs = v.g_isd = v.g_lc = nullptr;
    v.cpg = 0;
    for (int i = 0; i < NVAL; i++) v.val_defaults[i] = 0.0;
    if (plan) {
        v.val_defaults[NGRTD_P_F1] = 1.0;                                     // p_dict defaults, run_age_mcmc_utils.py:73-79
        v.val_defaults[NGRTD_P_J] = plan->pv.default_log10J;
    } else {
        v.gases.n = cfg->ngas;
        for (int g = 0; g < cfg->ngas; g++) {
            if (cfg->gases[g] < 0 || cfg->gases[g] > 4) { delete S; return fail(NGRTD_EINVAL, "sampler: gas id must be 0..4"); }
            v.gases.id[g] = cfg->gases[g];
        }
    }
    const size_t B = (size_t)nchains, nd = (size_t)cfg->ndim;
    S->n_q = B * nd;
    size_t hist_bytes = (size_t)cfg->hist_cap * B * nd * sizeof(double);
    size_t free_b = 0, total_b = 0;
    CUDA_TRY_OR(cudaMemGetInfo(&free_b, &total_b), delete S);
    if (hist_bytes > free_b / 2) {
        delete S;
        return fail(NGRTD_ENOMEM, "sampler: history ring (" + std::to_string(hist_bytes >> 20) + " MiB) exceeds half of free device memory; lower hist_cap");
    }
    cudaError_t e;
    if ((e = cudaMalloc((void**)&v.q, B * nd * 8)) || (e = cudaMalloc((void**)&v.logp, B * 8)) ||
        (e = cudaMalloc((void**)&v.lamb, B * 8)) || (e = cudaMalloc((void**)&v.scal, B * 8)) ||
        (e = cudaMalloc((void**)&v.acc_win, B * 4)) || (e = cudaMalloc((void**)&v.acc_tot, B * 8)) ||
        (e = cudaMalloc((void**)&v.hist, hist_bytes)) || (e = cudaMalloc((void**)&v.wf_mean, B * nd * 8)) ||
        (e = cudaMalloc((void**)&v.wf_m2, B * nd * 8))) {
        ngrtd_sampler_destroy(S);
        return fail(NGRTD_ENOMEM, std::string("sampler alloc: ") + cudaGetErrorString(e));
    }
    {
        std::vector<double> hq(B * nd), hl(B, cfg->lamb > 0 ? cfg->lamb : 2.38 / std::sqrt(2.0 * cfg->ndim)), hs(B, cfg->scaling);
        for (size_t b = 0; b < B; b++) for (size_t d = 0; d < nd; d++) hq[b * nd + d] = start[d];
        CUDA_TRY_OR(cudaMemcpy(v.q, hq.data(), B * nd * 8, cudaMemcpyHostToDevice), ngrtd_sampler_destroy(S));
        CUDA_TRY_OR(cudaMemcpy(v.lamb, hl.data(), B * 8, cudaMemcpyHostToDevice), ngrtd_sampler_destroy(S));
        CUDA_TRY_OR(cudaMemcpy(v.scal, hs.data(), B * 8, cudaMemcpyHostToDevice), ngrtd_sampler_destroy(S));
        CUDA_TRY_OR(cudaMemset(v.logp, 0, B * 8), ngrtd_sampler_destroy(S));
        CUDA_TRY_OR(cudaMemset(v.acc_win, 0, B * 4), ngrtd_sampler_destroy(S));
        CUDA_TRY_OR(cudaMemset(v.acc_tot, 0, B * 8), ngrtd_sampler_destroy(S));
        CUDA_TRY_OR(cudaMemset(v.wf_mean, 0, B * nd * 8), ngrtd_sampler_destroy(S));
        CUDA_TRY_OR(cudaMemset(v.wf_m2, 0, B * nd * 8), ngrtd_sampler_destroy(S));
    }
    RunArgs ra{};
    ra.mode = 1;                    // logp of the starting point
    ra.nsteps = 1;
    ra.thin = 1;
    int rc = sampler_launch(S, ra, nullptr);
    if (rc == NGRTD_OK) {
        cudaError_t e2 = cudaDeviceSynchronize();
        if (e2 != cudaSuccess) rc = fail(NGRTD_ECUDA, std::string("sampler init: ") + cudaGetErrorString(e2));
    }
    if (rc) { ngrtd_sampler_destroy(S); return rc; }
    *out = S;
    return NGRTD_OK;
}

// per-group observations: group g owns global chains [g*cpg, (g+1)*cpg)
extern "C" int ngrtd_sampler_set_obs_groups(ngrtd_sampler* S, const double* obs_mu, const double* obs_sd, int64_t ngroups,
                                            int64_t chains_per_group) {
    if (!S || !obs_mu || !obs_sd) return fail(NGRTD_EINVAL, "obs_groups: null pointer");
    DeviceGuard guard(S->device);
    if (!guard.ok()) return fail(NGRTD_ECUDA, "ngrtd_sampler_set_obs_groups: cudaSetDevice(sampler device) failed");
    if (ngroups < 1 || chains_per_group < 1) return fail(NGRTD_EINVAL, "obs_groups: need ngroups, chains_per_group >= 1");
    SamplerView& v = S->sv;
    if ((v.chain_offset + v.B + chains_per_group - 1) / chains_per_group > ngroups)
        return fail(NGRTD_EINVAL, "obs_groups: the shard's chains reach beyond the last group");
    const size_t n = (size_t)ngroups * v.ntr;
    std::vector<double> h(3 * n);
    for (size_t i = 0; i < n; i++) {
        double sd = obs_sd[i];
        h[i] = obs_mu[i];
        h[n + i] = 1.0 / sd;
        h[2 * n + i] = v.lik_kind == NGRTD_LIK_NORMAL ? -0.5 * std::log(2.0 * M_PI * sd * sd) : -std::log(sd);
    }
    cudaFree(S->d_groups);
    S->d_groups = nullptr;
    CUDA_TRY(cudaMalloc((void**)&S->d_groups, 3 * n * sizeof(double)));
    CUDA_TRY(cudaMemcpy(S->d_groups, h.data(), 3 * n * sizeof(double), cudaMemcpyHostToDevice));
    v.g_obs = S->d_groups;
    v.g_isd = S->d_groups + n;
    v.g_lc = S->d_groups + 2 * n;
    v.cpg = chains_per_group;
    RunArgs ra{};                    // the likelihood changed: refresh logp of the current state
    ra.mode = 1; ra.nsteps = 1; ra.thin = 1;
    int rc = sampler_launch(S, ra, nullptr);
    if (rc) return rc;
    CUDA_TRY(cudaDeviceSynchronize());
    return NGRTD_OK;
}

extern "C" int ngrtd_sampler_set_population(ngrtd_sampler* S, int64_t chains_per_population) {
    if (!S) return fail(NGRTD_EINVAL, "set_population: null sampler");
    if (chains_per_population < 0) return fail(NGRTD_EINVAL, "set_population: chains_per_population must be >= 0");
    S->sv.pool = chains_per_population;
    return NGRTD_OK;
}

extern "C" int ngrtd_sampler_run(ngrtd_sampler* S, int64_t nsteps, int32_t tune, int32_t record, int32_t thin,
                                 double* trace_d, void* stream) {
    if (!S) return fail(NGRTD_EINVAL, "null sampler");
    DeviceGuard guard(S->device);
    if (!guard.ok()) return fail(NGRTD_ECUDA, "ngrtd_sampler_run: cudaSetDevice(sampler device) failed");
    if (nsteps < 0 || nsteps > 2000000000LL) return fail(NGRTD_EINVAL, "sampler: bad nsteps");
    if (nsteps == 0) return NGRTD_OK;
    if (thin < 1) return fail(NGRTD_EINVAL, "sampler: thin must be >= 1");
    RunArgs ra{};
    ra.step0 = S->step;
    ra.nsteps = (int)nsteps;
    ra.mode = 0;
    ra.tune = tune ? 1 : 0;
    ra.hist_start = S->hist_start;
    ra.trace = record ? trace_d : nullptr;
    ra.thin = thin;
    ra.draw0 = S->ndraws;
    ra.record = record ? 1 : 0;
    int rc = sampler_launch(S, ra, (cudaStream_t)stream);
    if (rc) return rc;
    S->step += nsteps;
    if (record) S->ndraws += (nsteps + thin - 1) / thin;
    return NGRTD_OK;
}

extern "C" int ngrtd_sampler_stop_tuning(ngrtd_sampler* S) {
    if (!S) return fail(NGRTD_EINVAL, "null sampler");
    DeviceGuard guard(S->device);
    if (!guard.ok()) return fail(NGRTD_ECUDA, "ngrtd_sampler_stop_tuning: cudaSetDevice(sampler device) failed");
    long long it = S->step - S->hist_start;              // len(self._history)
    long long n_drop = (long long)(S->tune_drop_fraction * (double)it);
    S->hist_start += n_drop;
    return NGRTD_OK;
}

extern "C" int ngrtd_sampler_info(const ngrtd_sampler* S, int64_t* step, int64_t* ndraws, int64_t* hist_start) {
    if (!S) return fail(NGRTD_EINVAL, "null sampler");
    if (step) *step = S->step;
    if (ndraws) *ndraws = S->ndraws;
    if (hist_start) *hist_start = S->hist_start;
    return NGRTD_OK;
}

extern "C" int ngrtd_sampler_set_counters(ngrtd_sampler* S, int64_t step, int64_t ndraws, int64_t hist_start) {
    if (!S) return fail(NGRTD_EINVAL, "null sampler");
    DeviceGuard guard(S->device);
    if (!guard.ok()) return fail(NGRTD_ECUDA, "ngrtd_sampler_set_counters: cudaSetDevice(sampler device) failed");
    if (step < 0 || ndraws < 0 || hist_start < 0 || hist_start > step) return fail(NGRTD_EINVAL, "sampler_set_counters: bad counters");
    S->step = step;
    S->ndraws = ndraws;
    S->hist_start = hist_start;
    return NGRTD_OK;
}

__global__ void k_double_to_int(const double* a, long long n, int* oi, long long* ol) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) { if (oi) oi[i] = (int)a[i]; else ol[i] = (long long)a[i]; }
}

__global__ void k_int_to_double(const int* a, const long long* b, long long n, double* out) {
    long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i < n) out[i] = b ? (double)b[i] : (double)a[i];
}

extern "C" int ngrtd_sampler_get(ngrtd_sampler* S, int32_t what, double* out_d, void* stream) {
    if (!S || !out_d) return fail(NGRTD_EINVAL, "sampler_get: null pointer");
    DeviceGuard guard(S->device);
    if (!guard.ok()) return fail(NGRTD_ECUDA, "ngrtd_sampler_get: cudaSetDevice(sampler device) failed");
    cudaStream_t st = (cudaStream_t)stream;
    const SamplerView& v = S->sv;
    size_t B = (size_t)v.B;
    const double* src = nullptr;
    size_t n = B;
    switch (what) {
        case 0: src = v.q; n = S->n_q; break;
        case 1: src = v.logp; break;
        case 2: src = v.lamb; break;
        case 3: src = v.scal; break;
        case 4:
        case 8:
            k_int_to_double<<<(unsigned)((B + 255) / 256), 256, 0, st>>>(what == 8 ? v.acc_win : nullptr,
                                                                         what == 8 ? nullptr : v.acc_tot, (long long)B, out_d);
            CUDA_TRY(cudaGetLastError());
            return NGRTD_OK;
        case 5: src = v.wf_mean; n = S->n_q; break;
        case 6: src = v.wf_m2; n = S->n_q; break;
        case 7: src = v.hist; n = (size_t)v.hist_cap * S->n_q; break;
        default: return fail(NGRTD_EINVAL, "sampler_get: unknown selector");
    }
    CUDA_TRY(cudaMemcpyAsync(out_d, src, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    return NGRTD_OK;
}

// ---- K6: pooled moments on the device.  Stage 1: every block sums its slice of chains for the 3*nd quantities (fixed
// chain -> block -> thread assignment and a fixed tree, so the result is deterministic); stage 2: one block folds the
// block partials.  POOL_BLOCKS * 3 * ND_MAX doubles of workspace live in the sampler object.
constexpr int POOL_BLOCKS = 296, POOL_THREADS = 256;
__global__ void __launch_bounds__(POOL_THREADS) k_pool_stage1(const double* __restrict__ mean, const double* __restrict__ m2,
                                                              long long B, int nd, double* __restrict__ part) {
    __shared__ double red[POOL_THREADS / 32][3 * ND_MAX];
    double acc[3 * ND_MAX];
#pragma unroll
    for (int i = 0; i < 3 * ND_MAX; i++) acc[i] = 0.0;
    const long long per = (B + gridDim.x - 1) / gridDim.x;
    const long long c0 = (long long)blockIdx.x * per, c1 = min(B, c0 + per);
    for (long long c = c0 + threadIdx.x; c < c1; c += blockDim.x) {
#pragma unroll
        for (int d = 0; d < ND_MAX; d++) {
            if (d < nd) {
                const double m = mean[c * nd + d];
                acc[d] += m;
                acc[ND_MAX + d] = fma(m, m, acc[ND_MAX + d]);
                acc[2 * ND_MAX + d] += m2[c * nd + d];
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 3 * ND_MAX; i++) {
        double v = acc[i];
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][i] = v;
    }
    __syncthreads();
    if (threadIdx.x < 3 * ND_MAX) {
        double v = 0.0;
        for (int w = 0; w < POOL_THREADS / 32; w++) v += red[w][threadIdx.x];
        part[(size_t)blockIdx.x * 3 * ND_MAX + threadIdx.x] = v;
    }
}
__global__ void k_pool_stage2(const double* __restrict__ part, int nblocks, int nd, long long B, double* __restrict__ out) {
    const int i = threadIdx.x;           // one thread per (quantity, dimension)
    if (i < 3 * ND_MAX) {
        double v = 0.0;
        for (int b = 0; b < nblocks; b++) v += part[(size_t)b * 3 * ND_MAX + i];
        const int q = i / ND_MAX, d = i % ND_MAX;
        if (d < nd) out[q * nd + d] = v;
    }
    if (i == 0) out[3 * nd] = (double)B;
}

extern "C" int ngrtd_sampler_pooled_moments(ngrtd_sampler* S, double* out_d, void* stream) {
    if (!S || !out_d) return fail(NGRTD_EINVAL, "pooled_moments: null pointer");
    DeviceGuard guard(S->device);
    if (!guard.ok()) return fail(NGRTD_ECUDA, "pooled_moments: cudaSetDevice failed");
    cudaStream_t st = (cudaStream_t)stream;
    if (!S->d_pool) CUDA_TRY(cudaMalloc((void**)&S->d_pool, sizeof(double) * POOL_BLOCKS * 3 * ND_MAX));
    const SamplerView& v = S->sv;
    k_pool_stage1<<<POOL_BLOCKS, POOL_THREADS, 0, st>>>(v.wf_mean, v.wf_m2, v.B, v.nd, S->d_pool);
    k_pool_stage2<<<1, 32, 0, st>>>(S->d_pool, POOL_BLOCKS, v.nd, v.B, out_d);
    CUDA_TRY(cudaGetLastError());
    return NGRTD_OK;
}

extern "C" int ngrtd_sampler_set(ngrtd_sampler* S, int32_t what, const double* in_d, void* stream) {
    if (!S || !in_d) return fail(NGRTD_EINVAL, "sampler_set: null pointer");
    DeviceGuard guard(S->device);
    if (!guard.ok()) return fail(NGRTD_ECUDA, "ngrtd_sampler_set: cudaSetDevice(sampler device) failed");
    cudaStream_t st = (cudaStream_t)stream;
    SamplerView& v = S->sv;
    double* dst = nullptr;
    size_t n = (size_t)v.B;
    switch (what) {
        case 0: dst = v.q; n = S->n_q; break;
        case 1: dst = v.logp; break;
        case 2: dst = v.lamb; break;
        case 3: dst = v.scal; break;
        case 4:
        case 8:
            k_double_to_int<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(in_d, (long long)n, what == 8 ? v.acc_win : nullptr,
                                                                         what == 8 ? nullptr : v.acc_tot);
            CUDA_TRY(cudaGetLastError());
            return NGRTD_OK;
        case 5: dst = v.wf_mean; n = S->n_q; break;
        case 6: dst = v.wf_m2; n = S->n_q; break;
        case 7: dst = v.hist; n = (size_t)v.hist_cap * S->n_q; break;
        default: return fail(NGRTD_EINVAL, "sampler_set: unknown selector");
    }
    CUDA_TRY(cudaMemcpyAsync(dst, in_d, n * sizeof(double), cudaMemcpyDeviceToDevice, st));
    if (what == 0) {                 // new positions: refresh logp
        RunArgs ra{};
        ra.mode = 1; ra.nsteps = 1; ra.thin = 1;
        return sampler_launch(S, ra, st);
    }
    return NGRTD_OK;
}

extern "C" int ngrtd_philox4x32_10(const uint32_t ctr[4], const uint32_t key[2], uint32_t out[4]) {
    if (!ctr || !key || !out) return fail(NGRTD_EINVAL, "philox: null pointer");
    unsigned int* d = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d, 16));
    k_philox_kat<<<1, 1>>>(make_uint4(ctr[0], ctr[1], ctr[2], ctr[3]), make_uint2(key[0], key[1]), d);
    cudaError_t e = cudaMemcpy(out, d, 16, cudaMemcpyDeviceToHost);
    cudaFree(d);
    if (e != cudaSuccess) return fail(NGRTD_ECUDA, cudaGetErrorString(e));
    return NGRTD_OK;
}

// ---------------------------------------------------------------- FP64 peak probe (roofline denominator, measured live)
// MEASURED_PEAKS.json of the pool has no FP64 entry; bench.py calls this so that the roofline fraction it prints is taken
// against a number measured in the same process on the same GPU (tools/microbench/fp64_peak.cu is the stand-alone study).
namespace {
constexpr int PEAK_ITERS = 4096, PEAK_ACC = 16;
__global__ void __launch_bounds__(256) k_peak_dfma(double* out, double a, double b) {
    double acc[PEAK_ACC];
#pragma unroll
    for (int i = 0; i < PEAK_ACC; i++) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
        for (int i = 0; i < PEAK_ACC; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < PEAK_ACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void __launch_bounds__(256) k_peak_dmma(double* out, double a, double b) {
    double c0[8], c1[8];
#pragma unroll
    for (int i = 0; i < 8; i++) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
    for (int it = 0; it < PEAK_ITERS; it++) {
#pragma unroll
        for (int i = 0; i < 8; i++)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                         : "+d"(c0[i]), "+d"(c1[i]) : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 8; i++) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

extern "C" int ngrtd_fp64_peak_probe(int32_t device, int32_t kind, double* tflops_out) {
    if (!tflops_out) return fail(NGRTD_EINVAL, "peak probe: null pointer");
    if (kind != 0 && kind != 1) return fail(NGRTD_EINVAL, "peak probe: kind must be 0 (DFMA) or 1 (DMMA m8n8k4)");
    int dev = device;
    if (dev < 0) CUDA_TRY(cudaGetDevice(&dev));
    DeviceGuard guard(dev);
    if (!guard.ok()) return fail(NGRTD_ECUDA, "peak probe: cudaSetDevice failed");
    int nsm = 0;
    CUDA_TRY(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = nsm * 4, threads = 256;             // 4 CTAs of 8 warps per SM: 8 warps per sub-partition
    double* d = nullptr;
    CUDA_TRY(cudaMalloc((void**)&d, (size_t)blocks * threads * sizeof(double)));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 6; rep++) {                   // first repetition warms up
        cudaEventRecord(e0);
        if (kind == 0) k_peak_dfma<<<blocks, threads>>>(d, 1.0000001, 1e-9);
        else k_peak_dmma<<<blocks, threads>>>(d, 1.0000001, 1e-9);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0 && ms < best) best = ms;
    }
    cudaError_t e = cudaGetLastError();
    cudaEventDestroy(e0); cudaEventDestroy(e1); cudaFree(d);
    if (e != cudaSuccess) return fail(NGRTD_ECUDA, cudaGetErrorString(e));
    // flops: DFMA = 2 per lane-instruction; DMMA m8n8k4 = 8*8*4*2 = 512 per warp-instruction
    const double flops = kind == 0 ? 2.0 * PEAK_ACC * PEAK_ITERS * (double)blocks * threads
                                   : 512.0 * 8 * PEAK_ITERS * (double)blocks * (threads / 32);
    *tflops_out = flops / (best * 1e-3) / 1e12;
    return NGRTD_OK;
}
#endif  // NGRTD_HAS_API
