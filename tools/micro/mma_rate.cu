// Microbenchmark: sustained warp-level mma.sync rate per SM on sm_100a (the legacy tensor path a recurrent
// step kernel can issue from ordinary warps without TMEM), as a function of warps per SM (1 CTA/SM).
//   mode 0: mma.sync.m16n8k8 tf32, 4x2 register tile (8 independent accumulators), operands fixed in registers
//   mode 1: 3xTF32 (hi*hi + lo*hi + hi*lo) with the hi/lo split of both operands done in the loop (LOP + FADD)
//   mode 2: mma.sync.m16n8k16 bf16, same tile
// Output: MAC/clk/SM (fp32-equivalent MACs for mode 1, i.e. raw MMA MACs / 3).
#include <cstdio>
#include <cuda_runtime.h>
#include <cstdint>

__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
                 : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ void split(float x, uint32_t& hi, uint32_t& lo) {
    hi = __float_as_uint(x) & 0xffffe000u;
    lo = __float_as_uint(x - __uint_as_float(hi));
}

template <int MODE>
__global__ void k(float* out, int iters, float s0) {
    float c[4][2][4];
    for (int m = 0; m < 4; ++m) for (int n = 0; n < 2; ++n) for (int i = 0; i < 4; ++i) c[m][n][i] = 0.f;
    float af[4][4], bf[2][2];
    for (int m = 0; m < 4; ++m) for (int i = 0; i < 4; ++i) af[m][i] = s0 + threadIdx.x * 1e-3f + m + i * 0.25f;
    for (int n = 0; n < 2; ++n) for (int i = 0; i < 2; ++i) bf[n][i] = s0 * 0.5f + n + i * 0.125f;
    for (int it = 0; it < iters; ++it) {
        if (MODE == 1) {
            uint32_t ah[4][4], al[4][4], bh[2][2], bl[2][2];
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int i = 0; i < 4; ++i) split(af[m][i], ah[m][i], al[m][i]);
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int i = 0; i < 2; ++i) split(bf[n][i], bh[n][i], bl[n][i]);
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int n = 0; n < 2; ++n) {
                    mma_tf32(c[m][n], al[m], bh[n]);
                    mma_tf32(c[m][n], ah[m], bl[n]);
                    mma_tf32(c[m][n], ah[m], bh[n]);
                }
        } else {
            uint32_t a[4][4], b[2][2];
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int i = 0; i < 4; ++i) a[m][i] = __float_as_uint(af[m][i]);
#pragma unroll
            for (int n = 0; n < 2; ++n)
#pragma unroll
                for (int i = 0; i < 2; ++i) b[n][i] = __float_as_uint(bf[n][i]);
#pragma unroll
            for (int m = 0; m < 4; ++m)
#pragma unroll
                for (int n = 0; n < 2; ++n) {
                    if (MODE == 0) mma_tf32(c[m][n], a[m], b[n]); else mma_bf16(c[m][n], a[m], b[n]);
                }
        }
#pragma unroll
        for (int m = 0; m < 4; ++m) af[m][0] += 1e-6f;   // operands change every iteration
        bf[0][0] += 1e-6f;
    }
    float s = 0;
    for (int m = 0; m < 4; ++m) for (int n = 0; n < 2; ++n) for (int i = 0; i < 4; ++i) s += c[m][n][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// dependent-issue latency: CH independent accumulator chains per warp, operands fixed
template <int CH>
__global__ void kc(float* out, int iters, float s0) {
    float c[CH][4];
    for (int m = 0; m < CH; ++m) for (int i = 0; i < 4; ++i) c[m][i] = 0.f;
    uint32_t a[4], b[2];
    for (int i = 0; i < 4; ++i) a[i] = __float_as_uint(s0 + threadIdx.x * 1e-3f + i);
    for (int i = 0; i < 2; ++i) b[i] = __float_as_uint(s0 * 0.5f + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int m = 0; m < CH; ++m) mma_tf32(c[m], a, b);
    }
    float s = 0;
    for (int m = 0; m < CH; ++m) for (int i = 0; i < 4; ++i) s += c[m][i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int CH>
void sweep(float* out, int clk) {
    for (int warps : {4, 8}) {
        const int iters = 20000;
        cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
        for (int rep = 0; rep < 2; ++rep) {
            cudaEventRecord(a);
            kc<CH><<<148, warps * 32>>>(out, iters, 1.f);
            cudaEventRecord(b); cudaEventSynchronize(b);
        }
        float ms; cudaEventElapsedTime(&ms, a, b);
        const double cyc = ms * 1e-3 * clk * 1e3 / iters;     // cycles per loop iteration (= CH MMAs per warp)
        printf("chains/warp %2d warps/SM %d : %.1f cycles per round  -> %.1f cycles per MMA per SMSP-warp, %.0f MAC/clk/SM\n", CH, warps, cyc,
               cyc / CH, (double)warps * CH * 1024 / cyc);
    }
}

int main() {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    const char* names[3] = {"tf32 m16n8k8     ", "3xTF32 (split in loop)", "bf16 m16n8k16    "};
    for (int mode = 0; mode < 3; ++mode)
        for (int warps : {4, 8, 16}) {
            const int iters = 20000;
            cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
            for (int rep = 0; rep < 2; ++rep) {
                cudaEventRecord(a);
                if (mode == 0) k<0><<<148, warps * 32>>>(out, iters, 1.f);
                else if (mode == 1) k<1><<<148, warps * 32>>>(out, iters, 1.f);
                else k<2><<<148, warps * 32>>>(out, iters, 1.f);
                cudaEventRecord(b); cudaEventSynchronize(b);
            }
            float ms; cudaEventElapsedTime(&ms, a, b);
            // MACs per warp per iteration: 8 tiles x (16*8*K); fp32-equivalent for mode 1
            double per_tile = mode == 2 ? 16.0 * 8 * 16 : 16.0 * 8 * 8;
            double mac = (double)148 * warps * iters * 8 * per_tile;
            cudaError_t e = cudaGetLastError();
            printf("%s warps/SM %2d : %.3f ms  %.1f TFLOP/s-equiv  %.0f MAC/clk/SM (at %.2f GHz nominal)%s\n", names[mode], warps, ms,
                   2 * mac / ms / 1e9, mac / 148 / (ms * 1e-3) / (clk * 1e3), clk / 1e6, e ? cudaGetErrorString(e) : "");
        }
    sweep<1>(out, clk); sweep<2>(out, clk); sweep<3>(out, clk); sweep<4>(out, clk); sweep<6>(out, clk);
    sweep<8>(out, clk); sweep<12>(out, clk); sweep<16>(out, clk);
    return 0;
}
