// K6 / K9 / K11: the dense "x . W^T (+ bias)" contractions of the step on the 5th-generation tensor cores, fp32 in and
// out, fp32-class accuracy ("3xTF32"):   C[M,N] (+)= A[M,K] . B[N,K]^T + bias[N]     (both operands K-major)
//   reference call sites: the input projection inside nn.LSTM (src/module.py:112-113,131), the CTC head (src/asr.py:29,
//   96), proj_k / char_trans / pj (src/asr.py:242-243,177,220; src/module.py:123,155) and their input gradients.
//
// tcgen05.mma kind::tf32 reads fp32 bit patterns from shared memory and IGNORES the low 13 mantissa bits (truncation,
// verified on hardware by tools/micro/umma_probe.cu), so the raw fp32 tile IS the TF32 "hi" operand - no split pass,
// no second copy in HBM.  Only the residual  lo = x - trunc(x)  has to exist as its own tile, and it has the same
// shared-memory layout as the raw tile, so making it is a purely elementwise pass by 8 "splitter" warps.  Per 32-wide
// K block the single-thread issuer then accumulates   A.B + A_lo.B + A.B_lo   in one TMEM accumulator (fp32).
//
// One 128 x 256 tile per CTA.  Warp roles: 0 = TMA producer (cp.async.bulk.tensor, 128-byte swizzle, out-of-bounds
// rows / K tail zero-filled by the hardware), 1 = MMA issue + TMEM allocation, 2..9 = splitters; 2..5 also drain the
// accumulator (tcgen05.ld 32x32b -> + bias [-> + C] -> 128-bit stores).  Two shared-memory stages of 96 KB
// (A 16 + B 32 raw, the same again for the residuals).
#include <cuda.h>
#include "common.cuh"
#include "umma.cuh"
#include "../../include/b200asr.h"

namespace b200asr {
namespace {

constexpr int G_BM = 128, G_BN = 256, G_BK = 32;
constexpr int G_STAGES = 2;
constexpr int G_A_BYTES = G_BM * G_BK * 4;      // 16 KB
constexpr int G_B_BYTES = G_BN * G_BK * 4;      // 32 KB
constexpr int G_STAGE_BYTES = 2 * (G_A_BYTES + G_B_BYTES);
constexpr int G_SPLIT_WARPS = 8;
constexpr int G_THREADS = 32 * (2 + G_SPLIT_WARPS);

__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(smem_dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void g_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ float tf32_residual(float x) {
    return x - __uint_as_float(__float_as_uint(x) & 0xffffe000u);
}

__global__ void __launch_bounds__(G_THREADS, 1)
gemm3x_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                 const float* __restrict__ bias, float* __restrict__ C, int M, int N, int K, int ldc, int accumulate) {
    extern __shared__ __align__(1024) uint8_t smem[];
    // stage s: [A raw | B raw | A lo | B lo]
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + G_STAGES * G_STAGE_BYTES);   // TMA landed
    uint64_t* split = full + G_STAGES;                                               // residual tiles written
    uint64_t* empty = split + G_STAGES;                                              // MMAs of the stage retired
    uint64_t* acc_done = empty + G_STAGES;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_done + 1);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int m0 = blockIdx.y * G_BM, n0 = blockIdx.x * G_BN;
    const int KB = (K + G_BK - 1) / G_BK;

    if (tid == 0) {
        for (int s = 0; s < G_STAGES; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&split[s], G_SPLIT_WARPS);
            mbar_init(&empty[s], 1);
        }
        mbar_init(acc_done, 1);
        mbar_fence_init();
    }
    if (warp == 1) umma::tmem_alloc(tmem_slot, 256);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    if (warp == 0) {
        if (lane == 0) {
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % G_STAGES;
                if (kb >= G_STAGES) mbar_wait(&empty[s], (uint32_t)(((kb / G_STAGES) - 1) & 1));
                uint8_t* st = smem + s * G_STAGE_BYTES;
                mbar_expect_tx(&full[s], G_A_BYTES + G_B_BYTES);
                tma_load_2d(st, &map_a, kb * G_BK, m0, &full[s]);
                tma_load_2d(st + G_A_BYTES, &map_b, kb * G_BK, n0, &full[s]);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = umma::instr_desc(umma::FMT_TF32, G_BM, G_BN);
            for (int kb = 0; kb < KB; ++kb) {
                const int s = kb % G_STAGES;
                mbar_wait(&split[s], (uint32_t)((kb / G_STAGES) & 1));
                umma::fence_after_sync();
                const uint32_t a = smem_u32(smem + s * G_STAGE_BYTES), b = a + G_A_BYTES;
                const uint32_t alo = b + G_B_BYTES, blo = alo + G_A_BYTES;
#pragma unroll
                for (int k4 = 0; k4 < G_BK / 8; ++k4) {
                    const uint64_t da = umma::desc_k_sw128(a + k4 * 32), db = umma::desc_k_sw128(b + k4 * 32);
                    umma::mma_ss<umma::FMT_TF32>(tmem, da, db, idesc, (kb | k4) != 0);
                    umma::mma_ss<umma::FMT_TF32>(tmem, umma::desc_k_sw128(alo + k4 * 32), db, idesc, 1);
                    umma::mma_ss<umma::FMT_TF32>(tmem, da, umma::desc_k_sw128(blo + k4 * 32), idesc, 1);
                }
                umma::commit(&empty[s]);
            }
            umma::commit(acc_done);
        }
    } else {
        // ---------------------------------------------------------------- splitters: lo = x - trunc_tf32(x)
        const int st_tid = tid - 64;                                  // 0..255
        for (int kb = 0; kb < KB; ++kb) {
            const int s = kb % G_STAGES;
            mbar_wait(&full[s], (uint32_t)((kb / G_STAGES) & 1));
            const float4* src = reinterpret_cast<const float4*>(smem + s * G_STAGE_BYTES);
            float4* dst = reinterpret_cast<float4*>(smem + s * G_STAGE_BYTES + G_A_BYTES + G_B_BYTES);
#pragma unroll 4
            for (int i = st_tid; i < (G_A_BYTES + G_B_BYTES) / 16; i += 32 * G_SPLIT_WARPS) {
                const float4 v = src[i];
                dst[i] = make_float4(tf32_residual(v.x), tf32_residual(v.y), tf32_residual(v.z), tf32_residual(v.w));
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            __syncwarp();
            if (lane == 0) g_arrive(&split[s]);
        }
        // ---------------------------------------------------------------- epilogue (warps 2..5: one TMEM quadrant each)
        if (warp < 6) {
            const int q = warp & 3;
            mbar_wait(acc_done, 0);
            umma::fence_after_sync();
            const int row = m0 + 32 * q + lane;
            const bool vec = ((ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(C) & 15) == 0);
            for (int c0 = 0; c0 < G_BN; c0 += 32) {
                if (n0 + c0 >= N) break;
                uint32_t v[32];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    uint32_t t8[8];
                    umma::ld_32x32b_x8(tmem + ((uint32_t)(32 * q) << 16) + c0 + 8 * j, t8);
#pragma unroll
                    for (int i = 0; i < 8; ++i) v[8 * j + i] = t8[i];
                }
                umma::wait_ld();
                if (row < M) {
                    float* crow = C + (size_t)row * ldc + n0 + c0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int n = n0 + c0 + 4 * j;
                        float o[4];
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            o[i] = __uint_as_float(v[4 * j + i]);
                            if (bias && n + i < N) o[i] += bias[n + i];
                        }
                        if (vec && n + 3 < N) {
                            float4* p4 = reinterpret_cast<float4*>(crow + 4 * j);
                            if (accumulate) {
                                const float4 old = *p4;
                                o[0] += old.x; o[1] += old.y; o[2] += old.z; o[3] += old.w;
                            }
                            *p4 = make_float4(o[0], o[1], o[2], o[3]);
                        } else {
#pragma unroll
                            for (int i = 0; i < 4; ++i)
                                if (n + i < N) crow[4 * j + i] = accumulate ? crow[4 * j + i] + o[i] : o[i];
                        }
                    }
                }
            }
        }
    }
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 1) umma::tmem_dealloc(tmem, 256);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// row-major fp32 matrix [rows, K] with row pitch ld (floats; rows may overlap when ld < K: the im2col view of a
// strided convolution) -> tensor map with boxes of 32 (K) x box_rows, 128-byte swizzle
int make_map(CUtensorMap* map, const float* ptr, int rows, int K, int ld, int box_rows) {
    EncodeTiledFn enc = encode_fn();
    B200_REQUIRE(enc != nullptr, "gemm: cuTensorMapEncodeTiled is not available from this driver");
    const cuuint64_t dims[2] = {(cuuint64_t)K, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(float)};
    const cuuint32_t box[2] = {(cuuint32_t)G_BK, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    B200_REQUIRE(r == CUDA_SUCCESS, "gemm: cuTensorMapEncodeTiled failed (%d) for a [%d x %d] matrix", (int)r, rows, K);
    return B200_OK;
}

}  // namespace
}  // namespace b200asr

using namespace b200asr;

extern "C" int b200asr_gemm3x_supported(int M, int N, int K) {
    // TMA: 16-byte aligned row pitch; at least one full tile's worth of work is not required (tails are zero-filled)
    return (M > 0 && N > 0 && K > 0 && (K % 4) == 0) ? 1 : 0;
}

extern "C" int b200asr_gemm3x_tn(const float* A, const float* B, const float* bias, float* C, int M, int N, int K,
                                 int ldc, int accumulate, b200asr_stream stream) {
    return b200asr_gemm3x_tn_ld(A, K, B, bias, C, M, N, K, ldc, accumulate, stream);
}

extern "C" int b200asr_gemm3x_tn_ld(const float* A, int lda, const float* B, const float* bias, float* C, int M, int N,
                                    int K, int ldc, int accumulate, b200asr_stream stream) {
    B200_REQUIRE(A && B && C, "gemm3x_tn: null pointer");
    B200_REQUIRE(lda > 0 && (lda % 4) == 0, "gemm3x_tn: lda %d must be a positive multiple of 4", lda);
    B200_REQUIRE(b200asr_gemm3x_supported(M, N, K), "gemm3x_tn: unsupported sizes M=%d N=%d K=%d (K %% 4 must be 0)", M,
                 N, K);
    B200_REQUIRE(ldc >= N, "gemm3x_tn: ldc %d < N %d", ldc, N);
    B200_REQUIRE(((reinterpret_cast<uintptr_t>(A) | reinterpret_cast<uintptr_t>(B)) & 15) == 0,
                 "gemm3x_tn: operands must be 16-byte aligned");
    CUtensorMap ma, mb;
    int rc = make_map(&ma, A, M, K, lda, G_BM);
    if (rc != B200_OK) return rc;
    rc = make_map(&mb, B, N, K, K, G_BN);
    if (rc != B200_OK) return rc;
    const size_t smem = (size_t)G_STAGES * G_STAGE_BYTES + 256;
    B200_CUDA(cudaFuncSetAttribute(gemm3x_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    dim3 grid((N + G_BN - 1) / G_BN, (M + G_BM - 1) / G_BM);
    gemm3x_tn_kernel<<<grid, G_THREADS, smem, (cudaStream_t)stream>>>(ma, mb, bias, C, M, N, K, ldc, accumulate);
    B200_LAUNCH_CHECK("gemm3x_tn_kernel");
    return B200_OK;
}
