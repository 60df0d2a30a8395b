// FP64 pipe microbenchmark for B200 (sm_100a): DFMA vs DMMA (mma.sync f64) vs mixed vs exp().
// Purpose: pick the instruction class for the lag-axis reduction kernel and provide the
// measured FP64 roofline denominator (MEASURED_PEAKS.json has no FP64 entry).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include <math.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int ITERS = 4096;

template <int NACC>
__global__ void k_dfma(double* out, double a, double b) {
    double acc[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__device__ __forceinline__ void dmma884(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void dmma1688(double& c0, double& c1, double& c2, double& c3,
                                         double a0, double a1, double a2, double a3, double b0, double b1) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
                 : "+d"(c0), "+d"(c1), "+d"(c2), "+d"(c3) : "d"(a0), "d"(a1), "d"(a2), "d"(a3), "d"(b0), "d"(b1));
}
__device__ __forceinline__ void dmma16816(double& c0, double& c1, double& c2, double& c3,
                                          const double* a, const double* b) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                 : "+d"(c0), "+d"(c1), "+d"(c2), "+d"(c3)
                 : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                   "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
}

template <int NACC>
__global__ void k_dmma884(double* out, double a, double b) {
    double c0[NACC], c1[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma884(c0[i], c1[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c0[i] + c1[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dmma1688(double* out, double a, double b) {
    double c[NACC][4];
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma1688(c[i][0], c[i][1], c[i][2], c[i][3], a, b, a, b, b, a);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dmma16816(double* out, double a, double b) {
    double c[NACC][4];
    double av[8], bv[4];
#pragma unroll
    for (int i = 0; i < 8; i++) av[i] = a + i * 1e-9;
#pragma unroll
    for (int i = 0; i < 4; i++) bv[i] = b + i * 1e-9;
#pragma unroll
    for (int i = 0; i < NACC; i++) { c[i][0] = threadIdx.x * 1e-3 + i; c[i][1] = i; c[i][2] = 1; c[i][3] = 2; }
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) dmma16816(c[i][0], c[i][1], c[i][2], c[i][3], av, bv);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// mixed: per iteration NM dmma884 (256 FMA each/warp = 8 per lane) + NF dfma per lane
template <int NM, int NF>
__global__ void k_mixed(double* out, double a, double b) {
    double c0[NM], c1[NM], acc[NF];
#pragma unroll
    for (int i = 0; i < NM; i++) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
#pragma unroll
    for (int i = 0; i < NF; i++) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NM; i++) dmma884(c0[i], c1[i], a, b);
#pragma unroll
        for (int i = 0; i < NF; i++) acc[i] = fma(acc[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NM; i++) s += c0[i] + c1[i];
#pragma unroll
    for (int i = 0; i < NF; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DFMA + independent integer/FP32 work, to see if non-FP64 issue is free in the DFMA shadow
template <int NF, int NI>
__global__ void k_dfma_int(double* out, double a, double b, int m) {
    double acc[NF]; unsigned u[NI];
#pragma unroll
    for (int i = 0; i < NF; i++) acc[i] = threadIdx.x * 1e-3 + i;
#pragma unroll
    for (int i = 0; i < NI; i++) u[i] = threadIdx.x + i;
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NF; i++) acc[i] = fma(acc[i], a, b);
#pragma unroll
        for (int i = 0; i < NI; i++) u[i] = (u[i] ^ m) + (u[i] >> 3);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NF; i++) s += acc[i];
#pragma unroll
    for (int i = 0; i < NI; i++) s += u[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DFMA fed from shared-memory broadcast loads (the lag loop's operand pattern)
template <int NCOL>
__global__ void k_dfma_lds(double* out, double a, int L) {
    extern __shared__ double xs[];
    for (int i = threadIdx.x; i < L * NCOL; i += blockDim.x) xs[i] = 1.0 + i * 1e-6;
    __syncthreads();
    double acc[NCOL];
#pragma unroll
    for (int i = 0; i < NCOL; i++) acc[i] = 0;
    double w = a + threadIdx.x * 1e-6;
    for (int rep = 0; rep < ITERS / 64; rep++) {
        for (int k = 0; k < L; k++) {
            w = w * 0.9999;
#pragma unroll
            for (int i = 0; i < NCOL; i++) acc[i] = fma(w, xs[k * NCOL + i], acc[i]);
        }
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < NCOL; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void k_exp(double* out, double a) {
    double x = -a * (threadIdx.x + 1) * 1e-3;
    double s = 0;
    for (int it = 0; it < ITERS; it++) {
        s += exp(x);
        x -= 1e-4;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_div(double* out, double a) {
    double x = a * (threadIdx.x + 1);
    double s = 0;
    for (int it = 0; it < ITERS; it++) {
        s += 1.0 / x;
        x += 1e-4;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
__global__ void k_rsqrt(double* out, double a) {
    double x = a * (threadIdx.x + 1);
    double s = 0;
    for (int it = 0; it < ITERS; it++) {
        s += rsqrt(x);
        x += 1e-4;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <typename F>
float timeit(F launch, int reps = 5) {
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(); CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < reps; r++) {
        CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (ms < best) best = ms;
    }
    return best;
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    int clk; CK(cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0));
    printf("device %s SMs %d clock %d kHz smem/SM %zu regs/SM %d\n", p.name, p.multiProcessorCount, clk, p.sharedMemPerMultiprocessor, p.regsPerMultiprocessor);
    int nsm = p.multiProcessorCount;
    double* out; CK(cudaMalloc(&out, sizeof(double) * nsm * 16 * 1024));
    const double a = 0.999999, b = 1e-7;
    for (int bpsm = 1; bpsm <= 8; bpsm *= 2) {
        for (int threads : {128, 256}) {
            int blocks = nsm * bpsm;
            double lanes = (double)blocks * threads;
            float ms;
            ms = timeit([&] { k_dfma<8><<<blocks, threads>>>(out, a, b); });
            printf("dfma<8>      blk/SM %d thr %d: %.3f ms  %.2f TFLOP/s\n", bpsm, threads, ms, lanes * ITERS * 8 * 2 / ms / 1e9);
            ms = timeit([&] { k_dfma<16><<<blocks, threads>>>(out, a, b); });
            printf("dfma<16>     blk/SM %d thr %d: %.3f ms  %.2f TFLOP/s\n", bpsm, threads, ms, lanes * ITERS * 16 * 2 / ms / 1e9);
            ms = timeit([&] { k_dmma884<4><<<blocks, threads>>>(out, a, b); });
            printf("dmma884<4>   blk/SM %d thr %d: %.3f ms  %.2f TFLOP/s\n", bpsm, threads, ms, lanes / 32 * ITERS * 4 * 512.0 / ms / 1e9);
            ms = timeit([&] { k_dmma884<8><<<blocks, threads>>>(out, a, b); });
            printf("dmma884<8>   blk/SM %d thr %d: %.3f ms  %.2f TFLOP/s\n", bpsm, threads, ms, lanes / 32 * ITERS * 8 * 512.0 / ms / 1e9);
            ms = timeit([&] { k_dmma1688<4><<<blocks, threads>>>(out, a, b); });
            printf("dmma1688<4>  blk/SM %d thr %d: %.3f ms  %.2f TFLOP/s\n", bpsm, threads, ms, lanes / 32 * ITERS * 4 * 2048.0 / ms / 1e9);
            ms = timeit([&] { k_dmma16816<4><<<blocks, threads>>>(out, a, b); });
            printf("dmma16816<4> blk/SM %d thr %d: %.3f ms  %.2f TFLOP/s\n", bpsm, threads, ms, lanes / 32 * ITERS * 4 * 4096.0 / ms / 1e9);
            ms = timeit([&] { k_mixed<2, 16><<<blocks, threads>>>(out, a, b); });
            printf("mixed<2mma,16fma> blk/SM %d thr %d: %.3f ms  %.2f TFLOP/s total (mma %.2f + fma %.2f)\n", bpsm, threads, ms,
                   lanes * ITERS * (2 * 8 + 16) * 2 / ms / 1e9, lanes * ITERS * 16 * 2 / ms / 1e9, lanes * ITERS * 16 * 2 / ms / 1e9);
            ms = timeit([&] { k_mixed<4, 8><<<blocks, threads>>>(out, a, b); });
            printf("mixed<4mma,8fma>  blk/SM %d thr %d: %.3f ms  %.2f TFLOP/s total\n", bpsm, threads, ms,
                   lanes * ITERS * (4 * 8 + 8) * 2 / ms / 1e9);
            ms = timeit([&] { k_dfma_int<8, 8><<<blocks, threads>>>(out, a, b, 12345); });
            printf("dfma8+int8(24 alu ops) blk/SM %d thr %d: %.3f ms  %.2f TFLOP/s fp64\n", bpsm, threads, ms, lanes * ITERS * 8 * 2 / ms / 1e9);
            ms = timeit([&] { k_dfma_int<8, 3><<<blocks, threads>>>(out, a, b, 12345); });
            printf("dfma8+int3(9 alu ops)  blk/SM %d thr %d: %.3f ms  %.2f TFLOP/s fp64\n", bpsm, threads, ms, lanes * ITERS * 8 * 2 / ms / 1e9);
            CK(cudaFuncSetAttribute(k_dfma_lds<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024));
            ms = timeit([&] { k_dfma_lds<8><<<blocks, threads, 840 * 8 * 8>>>(out, a, 840); });
            printf("dfma_lds<8> L=840 blk/SM %d thr %d: %.3f ms  %.2f TFLOP/s (fma only; +1 dmul per 8)\n", bpsm, threads, ms,
                   lanes * (ITERS / 64) * 840.0 * 8 * 2 / ms / 1e9);
            ms = timeit([&] { k_exp<<<blocks, threads>>>(out, a); });
            printf("exp          blk/SM %d thr %d: %.3f ms  %.2f Gexp/s\n", bpsm, threads, ms, lanes * ITERS / ms / 1e6);
            ms = timeit([&] { k_div<<<blocks, threads>>>(out, a); });
            printf("div          blk/SM %d thr %d: %.3f ms  %.2f Gdiv/s\n", bpsm, threads, ms, lanes * ITERS / ms / 1e6);
            ms = timeit([&] { k_rsqrt<<<blocks, threads>>>(out, a); });
            printf("rsqrt        blk/SM %d thr %d: %.3f ms  %.2f Grsqrt/s\n", bpsm, threads, ms, lanes * ITERS / ms / 1e6);
        }
    }
    // sustained: DFMA for ~3 s to see the power-capped clock
    {
        int blocks = nsm * 4, threads = 256; double lanes = (double)blocks * threads;
        cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        CK(cudaEventRecord(e0));
        int n = 0;
        for (; n < 4000; n++) k_dfma<16><<<blocks, threads>>>(out, a, b);
        CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
        float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
        printf("sustained dfma<16> %d launches: %.1f ms  %.2f TFLOP/s\n", n, ms, lanes * ITERS * 16 * 2 * n / ms / 1e9);
    }
    return 0;
}
