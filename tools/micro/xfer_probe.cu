// How fast can 128 co-resident CTAs each pull the same 64 KB (per group of 32 CTAs) from L2 into shared memory?
// This is the all-gather of the next A operand in the tcgen05 LSTM forward kernel (csrc/lstm_umma.cu), isolated:
//   A  8 lanes (8 warps) x one 8 KB bulk copy           (what the kernel does today)
//   B  1 lane x 8 bulk copies of 8 KB
//   C  1 lane x 1 bulk copy of 64 KB
//   D  32 lanes x one 2 KB bulk copy
//   E  16 warps: LDG.128 (ld.relaxed.gpu) + STS.128
//   F  like A but every CTA reads a PRIVATE 64 KB buffer (no sharing of L2 lines between CTAs)
//   M  clusters of 8: each CTA issues ONE 8 KB bulk copy multicast to all 8 CTAs of its cluster
// All CTAs start each repetition together (global spin barrier); reported: cycles from issue to "all 64 KB landed",
// mean / max over CTAs, averaged over repetitions.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I end-to-end-asr-pytorch_b200/csrc
//        tools/micro/xfer_probe.cu -o tools/micro/xfer_probe
#include <cooperative_groups.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>
#include "common.cuh"

namespace cg = cooperative_groups;
using namespace b200asr;
namespace b200asr {
void set_error(const char*, ...) {}
void count_launch(int) {}
}  // namespace b200asr

#define CK(x)                                                                              \
    do {                                                                                   \
        cudaError_t e = (x);                                                               \
        if (e != cudaSuccess) {                                                            \
            printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(1);                                                                       \
        }                                                                                  \
    } while (0)

constexpr int BYTES = 65536, NCTA = 128, GROUP = 32, REPS = 200, THREADS = 32 * 18;

__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned target) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(ctr, 1u);
        while (*((volatile unsigned*)ctr) < target) {
        }
    }
    __syncthreads();
}
__device__ __forceinline__ uint4 ld_relaxed_v4(const void* p) {
    uint4 v;
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void bulk_g2s_mc(void* smem_dst, const void* gmem_src, uint32_t bytes, uint64_t* bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar)), "h"(mask)
        : "memory");
}

// mode: 0 A, 1 B, 2 C, 3 D, 4 E, 5 F, 6 M, 7 A', 8 E'
__global__ void __launch_bounds__(THREADS, 1) probe(const uint8_t* shared_buf, const uint8_t* private_buf, unsigned* ctr,
                                                    long long* out, int mode_in) {
    int mode = mode_in;
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ long long t_start;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int grp = blockIdx.x / GROUP;
    const uint8_t* src = mode == 5 ? private_buf + (size_t)blockIdx.x * BYTES : shared_buf + (size_t)grp * BYTES;
    if (tid == 0) {
        mbar_init(&bar, 1);
        mbar_fence_init();
    }
    __syncthreads();
    if (mode == 6) cg::this_cluster().sync();
    long long acc = 0, worst = 0;
    const bool rewrite = mode >= 7;                 // 7 = A, 8 = E, but every CTA REWRITES its 2 KB slice of the group buffer
    if (rewrite) mode = mode == 7 ? 0 : 4;          // before each repetition (freshly written lines, like the real exchange)
    for (int r = 0; r < REPS; ++r) {
        if (rewrite) {
            uint8_t* mine = const_cast<uint8_t*>(src) + (size_t)(blockIdx.x % GROUP) * 2048;
            if (tid < 128) {
                const uint32_t v = (uint32_t)r * 2654435761u + tid;
                asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%1,%1,%1};" ::"l"(mine + tid * 16), "r"(v) : "memory");
            }
        }
        grid_barrier(ctr, (unsigned)(r + 1) * gridDim.x);
        if (mode == 6) cg::this_cluster().sync();
        if (tid == 0) {
            t_start = clock64();
            if (mode != 4) mbar_expect_tx(&bar, BYTES);
        }
        __syncthreads();
        if (mode == 0 || mode == 5) {
            if (lane == 0 && warp < 8) bulk_g2s(smem + warp * 8192, src + warp * 8192, 8192, &bar);
        } else if (mode == 1) {
            if (tid == 0)
                for (int a = 0; a < 8; ++a) bulk_g2s(smem + a * 8192, src + a * 8192, 8192, &bar);
        } else if (mode == 2) {
            if (tid == 0) bulk_g2s(smem, src, BYTES, &bar);
        } else if (mode == 3) {
            if (warp == 0) bulk_g2s(smem + lane * 2048, src + lane * 2048, 2048, &bar);
        } else if (mode == 6) {
            const unsigned rank = cg::this_cluster().block_rank();
            if (tid == 0) bulk_g2s_mc(smem + rank * 8192, src + rank * 8192, 8192, &bar, (uint16_t)0xFF);
        } else if (mode == 4) {
            if (warp < 16) {
                uint4 v[8];
#pragma unroll
                for (int j = 0; j < 8; ++j) v[j] = ld_relaxed_v4(src + ((size_t)(j * 16 + warp) * 32 + lane) * 16);
#pragma unroll
                for (int j = 0; j < 8; ++j) *reinterpret_cast<uint4*>(smem + ((size_t)(j * 16 + warp) * 32 + lane) * 16) = v[j];
            }
        }
        if (mode == 4) {
            __syncthreads();
        } else if (tid == 0) {
            mbar_wait(&bar, (uint32_t)(r & 1));
        }
        if (tid == 0) {
            const long long dt = clock64() - t_start;
            acc += dt;
            if (dt > worst) worst = dt;
        }
        __syncthreads();
    }
    if (tid == 0) {
        out[blockIdx.x * 2] = acc / REPS;
        out[blockIdx.x * 2 + 1] = worst;
    }
}

int main() {
    uint8_t *sb, *pb;
    unsigned* ctr;
    long long* out;
    CK(cudaMalloc(&sb, 4 * BYTES));
    CK(cudaMalloc(&pb, (size_t)NCTA * BYTES));
    CK(cudaMalloc(&ctr, 4));
    CK(cudaMalloc(&out, NCTA * 2 * sizeof(long long)));
    CK(cudaMemset(sb, 1, 4 * BYTES));
    CK(cudaMemset(pb, 1, (size_t)NCTA * BYTES));
    const size_t smem = BYTES + 1024 + 96 * 1024;   // > half of the SM: one CTA per SM
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CK(cudaFuncSetAttribute(probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    const char* names[] = {"A 8 lanes x 8 KB bulk", "B 1 lane x 8 x 8 KB bulk", "C 1 lane x 64 KB bulk", "D 32 lanes x 2 KB bulk",
                           "E 16 warps LDG.128+STS", "F 8 x 8 KB bulk, private buffers", "M cluster-8 multicast 8 KB each",
                           "A' 8 x 8 KB bulk, slices rewritten", "E' LDG.128+STS, slices rewritten"};
    for (int mode = 0; mode < 9; ++mode) {
        CK(cudaMemset(ctr, 0, 4));
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(NCTA);
        cfg.blockDim = dim3(THREADS);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = mode == 6 ? 8 : 1;
        attr[0].val.clusterDim.y = 1;
        attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        if (mode == 6) {
            int ncl = 0;
            cudaError_t e = cudaOccupancyMaxActiveClusters(&ncl, probe, &cfg);
            printf("max active clusters of 8: %d (%s)\n", ncl, cudaGetErrorString(e));
            if (e != cudaSuccess || ncl * 8 < NCTA) { printf("%-34s skipped (clusters do not fit)\n", names[mode]); continue; }
        }
        const uint8_t* a0 = sb; const uint8_t* a1 = pb;
        CK(cudaLaunchKernelEx(&cfg, probe, a0, a1, ctr, out, mode));
        CK(cudaDeviceSynchronize());
        std::vector<long long> h(NCTA * 2);
        CK(cudaMemcpy(h.data(), out, h.size() * sizeof(long long), cudaMemcpyDeviceToHost));
        double mean = 0; long long mx = 0, mxmean = 0;
        for (int i = 0; i < NCTA; ++i) { mean += h[2 * i]; if (h[2 * i] > mxmean) mxmean = h[2 * i]; if (h[2 * i + 1] > mx) mx = h[2 * i + 1]; }
        printf("%-34s mean %6.0f cycles   slowest CTA (mean) %6lld   worst single %6lld   -> %5.1f B/clk/SM\n", names[mode],
               mean / NCTA, mxmean, mx, BYTES / (mean / NCTA));
    }
    return 0;
}
