// Does interleaving DMMA.8x8x4 with DFMA on B200's shared FP64 pipe cost switch bubbles?
// Patterns per loop iteration (per warp): FINE = 4 x [1 DMMA, 4 DFMA]; COARSE = [4 DMMA][16 DFMA]; COARSE2 = [8 DMMA][32 DFMA]
#include <cstdio>
#include <cuda_runtime.h>
constexpr int ITERS = 2048;
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
template <int MODE>
__global__ void k(double* out, double a, double b) {
    double c0[8], c1[8], acc[32];
    for (int i = 0; i < 8; i++) { c0[i] = threadIdx.x * 1e-3 + i; c1[i] = i; }
    for (int i = 0; i < 32; i++) acc[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < ITERS; it++) {
        if (MODE == 0) {        // fine
#pragma unroll
            for (int g = 0; g < 4; g++) {
                dmma(c0[g], c1[g], a, b);
#pragma unroll
                for (int i = 0; i < 4; i++) acc[g * 4 + i] = fma(acc[g * 4 + i], a, b);
            }
        } else if (MODE == 1) {  // coarse 4/16
#pragma unroll
            for (int g = 0; g < 4; g++) dmma(c0[g], c1[g], a, b);
#pragma unroll
            for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], a, b);
        } else if (MODE == 2) {  // coarse 8/32 (two iterations' worth)
#pragma unroll
            for (int g = 0; g < 8; g++) dmma(c0[g], c1[g], a, b);
#pragma unroll
            for (int i = 0; i < 32; i++) acc[i] = fma(acc[i], a, b);
        } else if (MODE == 3) {  // dmma only x4
#pragma unroll
            for (int g = 0; g < 4; g++) dmma(c0[g], c1[g], a, b);
        } else if (MODE == 4) {  // dfma only x16
#pragma unroll
            for (int i = 0; i < 16; i++) acc[i] = fma(acc[i], a, b);
        } else if (MODE == 5) {  // 1 DMMA : 1 DFMA fine
#pragma unroll
            for (int g = 0; g < 8; g++) { dmma(c0[g], c1[g], a, b); acc[g] = fma(acc[g], a, b); }
        }
        asm volatile("" ::: "memory");
    }
    double s = 0;
    for (int i = 0; i < 8; i++) s += c0[i] + c1[i];
    for (int i = 0; i < 32; i++) s += acc[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <typename F> float timeit(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize(); float best = 1e30f;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}
int main() {
    double* out; cudaMalloc(&out, 8 * 148 * 1024);
    const char* names[] = {"fine 4x[1 DMMA,4 DFMA]", "coarse [4 DMMA][16 DFMA]", "coarse [8 DMMA][32 DFMA]", "dmma only x4", "dfma only x16", "fine 8x[1 DMMA,1 DFMA]"};
    double ideal[] = {96, 96, 192, 64, 32, 144};
    for (int wps = 1; wps <= 8; wps *= 2) {
        int threads = 128 * wps > 1024 ? 1024 : 128 * wps;
        int blocks = 148 * (128 * wps / threads);
        for (int m = 0; m < 6; m++) {
            float ms = 0;
            switch (m) {
                case 0: ms = timeit([&] { k<0><<<blocks, threads>>>(out, 0.999, 1e-7); }); break;
                case 1: ms = timeit([&] { k<1><<<blocks, threads>>>(out, 0.999, 1e-7); }); break;
                case 2: ms = timeit([&] { k<2><<<blocks, threads>>>(out, 0.999, 1e-7); }); break;
                case 3: ms = timeit([&] { k<3><<<blocks, threads>>>(out, 0.999, 1e-7); }); break;
                case 4: ms = timeit([&] { k<4><<<blocks, threads>>>(out, 0.999, 1e-7); }); break;
                case 5: ms = timeit([&] { k<5><<<blocks, threads>>>(out, 0.999, 1e-7); }); break;
            }
            double cyc = ms * 1e-3 * 1.92e9 / ITERS / wps;
            printf("warps/SMSP %d  %-28s %.3f ms  %.1f cycles per warp-iteration (pipe-ideal %.0f)  eff %.0f%%\n", wps, names[m], ms, cyc, ideal[m], 100 * ideal[m] / cyc);
        }
    }
    return 0;
}
