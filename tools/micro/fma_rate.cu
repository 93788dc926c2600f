// Microbenchmark: sustained fp32 FMA rate per SM for scalar FFMA and packed FFMA2 register tiles, as a function of
// warps per SM (1 CTA/SM).  Answers: can 8 warps/SM (2 per scheduler) saturate the FMA pipes?
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned long long ffma2_vs(unsigned long long a, float b, unsigned long long c) {
    unsigned long long bb, d;
    asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(bb), "l"(c));
    return d;
}
template <int MODE>
__global__ void k(float* out, int iters, float s0, float s1) {
    float h[8], w[4];
    for (int i = 0; i < 8; ++i) h[i] = s0 + threadIdx.x * 1e-3f + i;
    for (int i = 0; i < 4; ++i) w[i] = s1 + i * 0.5f;
    if (MODE == 0) {
        float acc[8][4];
        for (int i = 0; i < 8; ++i) for (int q = 0; q < 4; ++q) acc[i][q] = 0.f;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[i][q] = fmaf(h[i], w[q], acc[i][q]);
#pragma unroll
            for (int i = 0; i < 8; ++i) h[i] += 1e-6f;   // keep the compiler honest
        }
        float s = 0; for (int i = 0; i < 8; ++i) for (int q = 0; q < 4; ++q) s += acc[i][q];
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    } else {
        unsigned long long acc[4][4], hp[4];
        for (int i = 0; i < 4; ++i) for (int q = 0; q < 4; ++q) acc[i][q] = 0ull;
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int i = 0; i < 4; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(hp[i]) : "f"(h[2 * i]), "f"(h[2 * i + 1]));
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[i][q] = ffma2_vs(hp[i], w[q], acc[i][q]);
#pragma unroll
            for (int i = 0; i < 8; ++i) h[i] += 1e-6f;
        }
        float s = 0;
        for (int i = 0; i < 4; ++i) for (int q = 0; q < 4; ++q) { float lo, hi; asm("mov.b64 {%0,%1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc[i][q])); s += lo + hi; }
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
}
int main() {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    for (int mode = 0; mode < 2; ++mode)
        for (int warps : {4, 8, 16, 32}) {
            const int iters = 20000;
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(a);
                if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, 1.f, 2.f); else k<1><<<148, warps * 32>>>(out, iters, 1.f, 2.f);
                cudaEventRecord(b); cudaEventSynchronize(b);
            }
            float ms; cudaEventElapsedTime(&ms, a, b);
            double fma = (double)148 * warps * 32 * iters * 32;
            printf("%s warps/SM %2d : %.3f ms  %.1f TFLOP/s  %.1f FMA/clk/SM (at %.2f GHz nominal)\n", mode ? "FFMA2" : "FFMA ", warps, ms,
                   2 * fma / ms / 1e9, fma / 148 / (ms * 1e-3) / (clk * 1e3), clk / 1e6);
        }
    return 0;
}
