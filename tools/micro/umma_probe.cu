// Hardware probe for the tcgen05 conventions in csrc/umma.cuh (run on a B200 through gpurun):
//   * K-major SWIZZLE_128B descriptors, multi-atom K, instruction descriptor, commit -> mbarrier
//   * kind::tf32 with RAW fp32 operands: does the tensor core truncate or round the low 13 mantissa bits?
//   * kind::f16
//   * where the rows of an M=64 / M=128 accumulator live in TMEM, and the register mapping of tcgen05.ld.16x256b
//   * issue-to-completion time of the MMA sequences the LSTM step kernels use
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I end-to-end-asr-pytorch_b200/csrc
//        tools/micro/umma_probe.cu -o tools/micro/umma_probe
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "umma.cuh"

using namespace b200asr;

namespace b200asr {
void set_error(const char*, ...) {}
void count_launch(int) {}
}  // namespace b200asr

#define CK(x)                                                                          \
    do {                                                                               \
        cudaError_t e = (x);                                                           \
        if (e != cudaSuccess) {                                                        \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                   \
        }                                                                              \
    } while (0)

// A [M,K], B [N,K] row-major in global (fp32 bit patterns for tf32, __half for f16).  D dump: [128 lanes][N].
template <int FMT, int M, int N, int K>
__global__ void __launch_bounds__(128, 1) probe_kernel(const void* Ag, const void* Bg, float* dump, uint32_t* frag,
                                                       long long* cyc, int reps) {
    constexpr int ES = (FMT == umma::FMT_TF32) ? 4 : 2;
    constexpr int KA = 128 / ES;          // elements per K atom
    constexpr int NA = K / KA;            // atoms along K
    constexpr int UK = 32 / ES;           // K per MMA
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* sA = smem;                                 // NA atoms of M x 128 B
    uint8_t* sB = smem + NA * M * 128;                  // NA atoms of N x 128 B
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < M * K; i += 128) {
        int r = i / K, k = i % K;
        uint32_t off = (k / KA) * (M * 128) + umma::sw128_offset(r, (k % KA) * ES);
        if (ES == 4) *reinterpret_cast<uint32_t*>(sA + off) = reinterpret_cast<const uint32_t*>(Ag)[i];
        else *reinterpret_cast<uint16_t*>(sA + off) = reinterpret_cast<const uint16_t*>(Ag)[i];
    }
    for (int i = tid; i < N * K; i += 128) {
        int r = i / K, k = i % K;
        uint32_t off = (k / KA) * (N * 128) + umma::sw128_offset(r, (k % KA) * ES);
        if (ES == 4) *reinterpret_cast<uint32_t*>(sB + off) = reinterpret_cast<const uint32_t*>(Bg)[i];
        else *reinterpret_cast<uint16_t*>(sB + off) = reinterpret_cast<const uint16_t*>(Bg)[i];
    }
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    if (warp == 0) umma::tmem_alloc(&tmem_slot, 512);
    fence_proxy_async();            // generic-proxy smem writes -> visible to the tensor core (async proxy)
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = tmem_slot;
    constexpr uint32_t idesc = umma::instr_desc(FMT, M, N);
    long long t0 = 0, t1 = 0;
    uint32_t parity = 0;
    for (int rep = 0; rep < reps; ++rep) {
        if (tid == 0) {
            t0 = clock64();
            for (int k = 0; k < K / UK; ++k) {
                const int atom = (k * UK) / KA, in = (k * UK) % KA;
                uint64_t da = umma::desc_k_sw128(smem_u32(sA + atom * M * 128) + in * ES);
                uint64_t db = umma::desc_k_sw128(smem_u32(sB + atom * N * 128) + in * ES);
                umma::mma_ss<FMT>(tmem, da, db, idesc, k > 0);
            }
            umma::commit(&bar);
        }
        mbar_wait(&bar, parity);
        parity ^= 1;
        if (tid == 0) {
            t1 = clock64();
            if (cyc) cyc[rep] = t1 - t0;
        }
    }
    umma::fence_after_sync();
    // dump all 128 lanes x N columns
    for (int c = 0; c < N; c += 8) {
        uint32_t v[8];
        umma::ld_32x32b_x8(tmem + ((uint32_t)(warp * 32) << 16) + c, v);
        umma::wait_ld();
        for (int i = 0; i < 8; ++i) dump[(warp * 32 + lane) * N + c + i] = __uint_as_float(v[i]);
    }
    // 16x256b fragments: x1 at columns 0..7 and x4 at columns 0..31
    {
        uint32_t v[4];
        umma::ld_16x256b_x1(tmem + ((uint32_t)(warp * 32) << 16), v);
        umma::wait_ld();
        for (int i = 0; i < 4; ++i) frag[tid * 20 + i] = v[i];
        if (N >= 32) {
            uint32_t w[16];
            umma::ld_16x256b_x4(tmem + ((uint32_t)(warp * 32) << 16), w);
            umma::wait_ld();
            for (int i = 0; i < 16; ++i) frag[tid * 20 + 4 + i] = w[i];
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

static float tf32_trunc(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u &= 0xffffe000u;
    memcpy(&x, &u, 4);
    return x;
}
static float tf32_rna(float x) {
    uint32_t u;
    memcpy(&u, &x, 4);
    u += 0x1000u;
    u &= 0xffffe000u;
    memcpy(&x, &u, 4);
    return x;
}

template <int FMT, int M, int N, int K>
static void run(const char* name, bool coded) {
    constexpr int ES = (FMT == umma::FMT_TF32) ? 4 : 2;
    std::vector<float> A(M * K), B(N * K);
    srand(1);
    for (auto& v : A) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    for (auto& v : B) v = (rand() / (float)RAND_MAX - 0.5f) * 2.f;
    if (coded) {
        std::fill(A.begin(), A.end(), 0.f);
        std::fill(B.begin(), B.end(), 0.f);
        for (int r = 0; r < M; ++r) A[r * K + 0] = (float)r, A[r * K + K - 1] = 1.f;   // uses first and last K atom
        for (int c = 0; c < N; ++c) B[c * K + 0] = 1000.f, B[c * K + K - 1] = (float)c;
    }
    std::vector<__half> Ah(M * K), Bh(N * K);
    for (int i = 0; i < M * K; ++i) Ah[i] = __float2half(A[i]);
    for (int i = 0; i < N * K; ++i) Bh[i] = __float2half(B[i]);
    void *dA, *dB;
    float* dD;
    uint32_t* dF;
    long long* dC;
    CK(cudaMalloc(&dA, M * K * 4));
    CK(cudaMalloc(&dB, N * K * 4));
    CK(cudaMalloc(&dD, 128 * N * 4));
    CK(cudaMalloc(&dF, 128 * 20 * 4));
    CK(cudaMalloc(&dC, 64 * 8));
    CK(cudaMemset(dD, 0xff, 128 * N * 4));
    CK(cudaMemset(dF, 0, 128 * 20 * 4));
    if (ES == 4) {
        CK(cudaMemcpy(dA, A.data(), M * K * 4, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, B.data(), N * K * 4, cudaMemcpyHostToDevice));
    } else {
        CK(cudaMemcpy(dA, Ah.data(), M * K * 2, cudaMemcpyHostToDevice));
        CK(cudaMemcpy(dB, Bh.data(), N * K * 2, cudaMemcpyHostToDevice));
    }
    size_t smem = (size_t)(K * ES / 128) * (M + N) * 128;
    CK(cudaFuncSetAttribute(probe_kernel<FMT, M, N, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const int reps = 8;
    probe_kernel<FMT, M, N, K><<<1, 128, smem>>>(dA, dB, dD, dF, dC, reps);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    std::vector<float> D(128 * N);
    std::vector<uint32_t> F(128 * 20);
    std::vector<long long> C(reps);
    CK(cudaMemcpy(D.data(), dD, 128 * N * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(F.data(), dF, 128 * 20 * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(C.data(), dC, reps * 8, cudaMemcpyDeviceToHost));
    // reference
    auto lane_of = [&](int r) { return M == 128 ? r : (r / 16) * 32 + r % 16; };
    double e_exact = 0, e_trunc = 0, e_rna = 0, scale = 0;
    for (int r = 0; r < M; ++r)
        for (int c = 0; c < N; ++c) {
            double s0 = 0, s1 = 0, s2 = 0;
            for (int k = 0; k < K; ++k) {
                float a = A[r * K + k], b = B[c * K + k];
                if (ES == 2) a = __half2float(Ah[r * K + k]), b = __half2float(Bh[c * K + k]);
                s0 += (double)a * b;
                s1 += (double)tf32_trunc(a) * tf32_trunc(b);
                s2 += (double)tf32_rna(a) * tf32_rna(b);
            }
            double d = D[lane_of(r) * N + c];
            e_exact = fmax(e_exact, fabs(d - s0));
            e_trunc = fmax(e_trunc, fabs(d - s1));
            e_rna = fmax(e_rna, fabs(d - s2));
            scale = fmax(scale, fabs(s0));
        }
    printf("[%s] FMT=%d M=%d N=%d K=%d  max|D-ref|/scale: exact-inputs %.3e  trunc-tf32 %.3e  rna-tf32 %.3e   cycles/rep:",
           name, FMT, M, N, K, e_exact / scale, e_trunc / scale, e_rna / scale);
    for (int i = 0; i < reps; ++i) printf(" %lld", C[i]);
    printf("  (%d MMAs)\n", K / (32 / ES));
    if (coded) {
        // where do rows live?  D[r][c] = 1000 r + c
        printf("   lanes holding data (lane: row decoded from col 0):");
        for (int l = 0; l < 128; ++l) {
            float v = D[l * N + 0];
            if (v == v && fabsf(v) < 1e9f && (l < 4 || (l % 16) == 0 || (l % 16) == 15)) printf(" %d:%g", l, v / 1000.f);
        }
        printf("\n   16x256b.x1 warp0: thread t -> (row,col) of v0..v3:");
        for (int t = 0; t < 32; t += 1) {
            if (t % 8 == 0) printf("\n     ");
            printf(" t%d:", t);
            for (int i = 0; i < 4; ++i) {
                float v;
                memcpy(&v, &F[t * 20 + i], 4);
                int r = (int)(v / 1000.f), c = (int)(v - 1000.f * r);
                printf("(%d,%d)", r, c);
            }
        }
        printf("\n   16x256b.x4 warp1 thread 5 -> ");
        for (int i = 0; i < 16; ++i) {
            float v;
            memcpy(&v, &F[(32 + 5) * 20 + 4 + i], 4);
            int r = (int)(v / 1000.f), c = (int)(v - 1000.f * r);
            printf("(%d,%d)", r, c);
        }
        printf("\n");
    }
    cudaFree(dA), cudaFree(dB), cudaFree(dD), cudaFree(dF), cudaFree(dC);
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    printf("device %s sm_%d%d, %d SMs, clock %d kHz\n", p.name, p.major, p.minor, p.multiProcessorCount, p.clockRate);
    run<umma::FMT_TF32, 128, 64, 64>("tf32 coded", true);
    run<umma::FMT_TF32, 64, 32, 64>("tf32 coded", true);
    run<umma::FMT_TF32, 128, 64, 64>("tf32 random", false);
    run<umma::FMT_TF32, 64, 32, 64>("tf32 random", false);
    run<umma::FMT_TF32, 128, 256, 32>("tf32 random", false);
    run<umma::FMT_F16, 64, 32, 128>("f16 coded", true);
    run<umma::FMT_F16, 64, 32, 128>("f16 random", false);
    run<umma::FMT_F16, 128, 128, 128>("f16 random", false);
    // timing of the LSTM-step-shaped sequences (K = 512 per product, x3 products emulated by 3x K)
    run<umma::FMT_F16, 64, 32, 512>("f16 lstm-fwd 1 product", false);
    run<umma::FMT_F16, 64, 32, 1024>("f16 lstm-fwd 2 products", false);
    run<umma::FMT_F16, 128, 32, 512>("f16 lstm-fwd M128", false);
    run<umma::FMT_F16, 64, 64, 768>("f16 N64", false);
    run<umma::FMT_F16, 64, 256, 192>("f16 lstm-bwd N256 K64x3", false);
    run<umma::FMT_F16, 64, 128, 512>("f16 lstm M64 N128 K512", false);
    run<umma::FMT_F16, 128, 64, 512>("f16 lstm M128 N64 K512", false);
    run<umma::FMT_F16, 128, 48, 512>("f16 lstm M128 N48 K512", false);
    run<umma::FMT_F16, 64, 96, 640>("f16 lstm M64 N96 K640", false);
    run<umma::FMT_TF32, 64, 32, 512>("tf32 K512", false);
    run<umma::FMT_TF32, 128, 128, 192>("tf32 gemm tile", false);
    run<umma::FMT_TF32, 128, 256, 128>("tf32 gemm tile N256", false);
    return 0;
}
