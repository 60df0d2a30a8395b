// Progressive reconstruction of the exponential-class inner loop to find what caps DMMA issue on B200.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
// V0: DMMA with constant A,B.  V1: A varies (DMUL recurrence).  V2: + masked select (ISETP/FSEL).  V3: + B from LDS.64
// V4: V3 with B via LDS but no select. NT tiles per warp.
template <int V, int NT>
__global__ void __launch_bounds__(512, 1) k(double* out, double r4, int k0, int ngroups, int reps) {
    extern __shared__ double xs[];
    for (int i = threadIdx.x; i < 840 * 8; i += blockDim.x) xs[i] = 1.0 + i * 1e-6;
    __syncthreads();
    const int lane = threadIdx.x & 31, j = lane & 3, r = lane >> 2;
    double c0[NT], c1[NT], v[NT], rr[NT]; int kk[NT];
    for (int t = 0; t < NT; t++) { c0[t] = 0; c1[t] = 0; v[t] = 1.0 + lane * 1e-3 + t; rr[t] = r4 + t * 1e-9; kk[t] = k0 + t + (lane >> 3); }
    double bconst = 1.0 + lane;
    for (int rep = 0; rep < reps; rep++) {
        const double* pf = xs + j * 8 + r;
        int kcur = j;
#pragma unroll 2
        for (int g = 0; g < ngroups; g++) {
            double b = (V >= 3) ? pf[0] : bconst;
            pf += 32; kcur += 4;
#pragma unroll
            for (int t = 0; t < NT; t++) {
                double a = (V == 0) ? bconst : v[t];
                if (V == 2 || V == 3) a = (kcur >= kk[t]) ? v[t] : 0.0;
                dmma(c0[t], c1[t], a, b);
                if (V >= 1) v[t] *= rr[t];
            }
        }
    }
    double s = 0;
    for (int t = 0; t < NT; t++) s += c0[t] + c1[t] + v[t];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize(); float best = 1e30f;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}
template <int V, int NT> void run(double* out, int threads, const char* name) {
    int ngroups = 210, reps = 8;
    cudaFuncSetAttribute(k<V, NT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 840 * 64);
    float ms = timeit([&] { k<V, NT><<<148, threads, 840 * 64>>>(out, 0.9999, 3, ngroups, reps); });
    double wps = threads / 128.0;
    double cyc = ms * 1e-3 * 1.92e9 / (reps * ngroups * wps * NT);
    printf("%-44s NT=%d warps/SMSP=%.0f  %.3f ms  %.1f cycles per tile-group per SMSP (ideal %d)\n", name, NT, wps, ms, cyc, V == 0 ? 16 : 18);
}
int main() {
    double* out; cudaMalloc(&out, 8 * 148 * 512);
    for (int threads : {128, 256, 512}) {
        run<0, 2>(out, threads, "V0 dmma const operands");
        run<1, 2>(out, threads, "V1 + A from DMUL recurrence");
        run<2, 2>(out, threads, "V2 + mask select");
        run<3, 2>(out, threads, "V3 + B from LDS.64");
        run<4, 2>(out, threads, "V4 B from LDS.64, no select");
        run<3, 1>(out, threads, "V3 NT=1");
        run<3, 4>(out, threads, "V3 NT=4");
    }
    return 0;
}
