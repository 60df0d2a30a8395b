// Cost anatomy of the dispersion-class weight (table exp) on B200: cycles per weight per SMSP, 4 warps/SMSP,
// 4 independent weights per loop iteration per warp.  Each variant removes one ingredient.
#include <cstdio>
#include <cuda_runtime.h>
constexpr double LN2 = 0.693147180559945309417232121458;
constexpr double K = 32 / LN2, C1 = LN2 / 32, C2 = C1 * C1 / 2, C3 = C1 * C1 * C1 / 6, C4 = C1 * C1 * C1 * C1 / 24;
enum { FULL = 0, NO_CONV, NO_TABLE, NO_FLUSH, NO_EXPINS, POLY_ONLY, NO_POLY, CONV_ONLY, FULL_PLUS_DMMA, POLY2, MAGIC, MAGIC_CHK, MAGIC_DMMA, POLY3 };
#define OPAQUE2(a, b) asm volatile("" : "+d"(a), "+d"(b))
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int V>
__device__ __forceinline__ double weight(double ep, const unsigned* th) {
    int n; double r;
    if (V == NO_CONV || V == POLY_ONLY) { n = __double2loint(ep) >> 3; r = ep - 3.0; }
    else if (V == MAGIC || V == MAGIC_CHK) {
        double tm = ep + 6755399441055744.0;
        n = __double2loint(tm);
        r = ep - (tm - 6755399441055744.0);
        if (V == MAGIC_CHK) { unsigned d = (unsigned)__double2hiint(tm) - 0x4337ffffu; n = d <= 1u ? n : (int)0x80000000; }
    }
    else { n = __double2int_rn(ep); r = ep - __int2double_rn(n); }
    if (V == CONV_ONLY) return r + (n & 1);
    double p;
    if (V == NO_POLY) p = r;
    else if (V == POLY2) p = fma(r, fma(r, C2, C1), 1.0);
    else if (V == POLY3) p = fma(r, fma(r, fma(r, C3, C2), C1), 1.0);
    else p = fma(r, fma(r, fma(r, fma(r, C4, C3), C2), C1), 1.0);
    if (V == POLY_ONLY) return p * ep;
    int jx = n & 31;
    int hi, lo;
    if (V == NO_TABLE) { hi = 0x3ff00000; lo = jx; } else { hi = (int)th[jx]; lo = (int)th[32 + jx]; }
    if (V != NO_EXPINS) hi += (n >> 5) << 20;
    double w = __hiloint2double(hi, lo) * p;
    if (V != NO_FLUSH) w = (n < -1022 * 32) ? 0.0 : w;
    return w;
}
template <int V>
__global__ void __launch_bounds__(512, 1) k(double* out, double ap0, double bp0, int ngroups, int reps) {
    __shared__ __align__(128) unsigned th[64];
    __shared__ double itp[840];
    __shared__ double xs[840 * 4];
    if (threadIdx.x < 64) th[threadIdx.x] = 0x3ff00000u + threadIdx.x * 1000;
    for (int i = threadIdx.x; i < 840; i += blockDim.x) itp[i] = 1.0 / (i + 1);
    for (int i = threadIdx.x; i < 840 * 4; i += blockDim.x) xs[i] = 1.0 + i * 1e-6;
    __syncthreads();
    const int lane = threadIdx.x & 31, j = lane & 3, r = lane >> 2;
    double acc[2] = {0, 0}, c0[2] = {0, 0}, c1[2] = {0, 0}, u[2], ap[2], bp[2];
    for (int t = 0; t < 2; t++) { ap[t] = ap0 * (1 + lane * 1e-3 + t); bp[t] = bp0 * (1 + t * 1e-2); u[t] = 10.0 + lane; }
    for (int rep = 0; rep < reps; rep++) {
        const double* pi = itp + j;
        const double* pf = xs + j * 4 + (r & 3);
#pragma unroll 2
        for (int g = 0; g < ngroups; g++) {
            double it = pi[0]; pi += 4;
            double b = (V == FULL_PLUS_DMMA || V == MAGIC_DMMA) ? pf[0] : 0.0; pf += 16;
#pragma unroll
            for (int t = 0; t < 2; t++) {
                double ep = fma(ap[t], it, u[t]);
                double w = weight<V == FULL_PLUS_DMMA ? FULL : (V == MAGIC_DMMA ? MAGIC_CHK : V)>(ep, th);
                u[t] = fma(bp[t], 4.0, u[t]);
                if (V == FULL_PLUS_DMMA || V == MAGIC_DMMA) dmma(c0[t], c1[t], w, b); else acc[t] += w;
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc[0] + acc[1] + c0[0] + c1[0] + c0[1] + c1[1];
}
// burst structure: GB groups of weights first (FP64 + conversions), then GB*2 DMMAs back to back
template <int GB, int UA>
__global__ void __launch_bounds__(512, 1) kburst(double* out, double ap0, double bp0, int ngroups, int reps) {
    __shared__ __align__(128) unsigned th[64];
    __shared__ double itp[840];
    __shared__ double xs[840 * 4];
    if (threadIdx.x < 64) th[threadIdx.x] = 0x3ff00000u + threadIdx.x * 1000;
    for (int i = threadIdx.x; i < 840; i += blockDim.x) itp[i] = 1.0 / (i + 1);
    for (int i = threadIdx.x; i < 840 * 4; i += blockDim.x) xs[i] = 1.0 + i * 1e-6;
    __syncthreads();
    const int lane = threadIdx.x & 31, j = lane & 3, r = lane >> 2;
    double c0[2][UA], c1[2][UA], u[2], ap[2], bp[2];
    for (int t = 0; t < 2; t++) { ap[t] = ap0 * (1 + lane * 1e-3 + t); bp[t] = bp0 * (1 + t * 1e-2); u[t] = 10.0 + lane;
        for (int a = 0; a < UA; a++) { c0[t][a] = 0; c1[t][a] = 0; } }
    for (int rep = 0; rep < reps; rep++) {
        const double* pi = itp + j;
        const double* pf = xs + j * 4 + (r & 3);
        for (int g = 0; g + GB <= ngroups; g += GB) {
            double w[GB][2], b[GB];
#pragma unroll
            for (int q = 0; q < GB; q++) {
                double it = pi[0]; pi += 4;
                b[q] = pf[0]; pf += 16;
#pragma unroll
                for (int t = 0; t < 2; t++) {
                    double ep = fma(ap[t], it, u[t]);
                    w[q][t] = weight<FULL>(ep, th);
                    u[t] = fma(bp[t], 4.0, u[t]);
                }
            }
#pragma unroll
            for (int q = 0; q < GB; q++) OPAQUE2(w[q][0], w[q][1]);
            OPAQUE2(u[0], u[1]);
#pragma unroll
            for (int q = 0; q < GB; q++)
#pragma unroll
                for (int t = 0; t < 2; t++) dmma(c0[t][q % UA], c1[t][q % UA], w[q][t], b[q]);
#pragma unroll
            for (int a = 0; a < UA; a++) { OPAQUE2(c0[0][a], c1[0][a]); OPAQUE2(c0[1][a], c1[1][a]); }
            OPAQUE2(u[0], u[1]);
        }
    }
    double s = 0;
    for (int t = 0; t < 2; t++) for (int a = 0; a < UA; a++) s += c0[t][a] + c1[t][a];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize(); float best = 1e30f;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}
template <int V> void run(double* out, int threads, const char* name) {
    int ngroups = 210, reps = 8;
    float ms = timeit([&] { k<V><<<148, threads>>>(out, -30.0, -0.01, ngroups, reps); });
    double wps = threads / 128.0;
    printf("warps/SMSP=%.0f  %-40s %.3f ms  %.1f cycles per weight per SMSP\n", wps, name, ms, ms * 1e-3 * 1.92e9 / (reps * ngroups * wps * 2));
}
int main() {
    double* out; cudaMalloc(&out, 8 * 148 * 512);
    for (int threads : {256, 512}) {
        run<FULL>(out, threads, "FULL (2 DFMA + conv + poly4 + table + ins + flush + DADD)");
        run<NO_CONV>(out, threads, "no F2I/I2F");
        run<NO_TABLE>(out, threads, "no table gather");
        run<NO_FLUSH>(out, threads, "no flush select");
        run<NO_EXPINS>(out, threads, "no exponent insert");
        run<POLY2>(out, threads, "degree-2 polynomial");
        run<NO_POLY>(out, threads, "no polynomial");
        run<POLY_ONLY>(out, threads, "poly only (2+4 DFMA, DADD, DMUL, DADD)");
        run<CONV_ONLY>(out, threads, "conv only (2 DFMA, F2I, I2F, 3 DADD)");
        run<FULL_PLUS_DMMA>(out, threads, "FULL + DMMA");
        run<MAGIC>(out, threads, "magic rounding (3 DADD, no conv, no check)");
        run<MAGIC_CHK>(out, threads, "magic rounding + range check");
        run<MAGIC_DMMA>(out, threads, "magic+check + DMMA");
        run<POLY3>(out, threads, "degree-3 polynomial (conv)");
        {
            int ngroups = 208, reps = 8; double wps = threads / 128.0;
            float ms = timeit([&] { kburst<2, 1><<<148, threads>>>(out, -30.0, -0.01, ngroups, reps); });
            printf("warps/SMSP=%.0f  burst GB=2 UA=1 FULL + DMMA: %.3f ms  %.1f cycles per weight per SMSP\n", wps, ms, ms * 1e-3 * 1.92e9 / (reps * ngroups * wps * 2));
            ms = timeit([&] { kburst<4, 1><<<148, threads>>>(out, -30.0, -0.01, ngroups, reps); });
            printf("warps/SMSP=%.0f  burst GB=4 UA=1 FULL + DMMA: %.3f ms  %.1f cycles per weight per SMSP\n", wps, ms, ms * 1e-3 * 1.92e9 / (reps * ngroups * wps * 2));
            ms = timeit([&] { kburst<4, 2><<<148, threads>>>(out, -30.0, -0.01, ngroups, reps); });
            printf("warps/SMSP=%.0f  burst GB=4 UA=2 FULL + DMMA: %.3f ms  %.1f cycles per weight per SMSP\n", wps, ms, ms * 1e-3 * 1.92e9 / (reps * ngroups * wps * 2));
            ms = timeit([&] { kburst<8, 2><<<148, threads>>>(out, -30.0, -0.01, ngroups, reps); });
            printf("warps/SMSP=%.0f  burst GB=8 UA=2 FULL + DMMA: %.3f ms  %.1f cycles per weight per SMSP\n", wps, ms, ms * 1e-3 * 1.92e9 / (reps * ngroups * wps * 2));
            ms = timeit([&] { kburst<8, 4><<<148, threads>>>(out, -30.0, -0.01, ngroups, reps); });
            printf("warps/SMSP=%.0f  burst GB=8 UA=4 FULL + DMMA: %.3f ms  %.1f cycles per weight per SMSP\n", wps, ms, ms * 1e-3 * 1.92e9 / (reps * ngroups * wps * 2));
        }
    }
    return 0;
}
