// K7/K8: persistent (Bi)LSTM recurrence, forward and BPTT backward, fp32.
//
// Decomposition (same for fwd and bwd): one CTA per (direction, batch-group, unit-block). A CTA owns
// UB hidden units (= 4*UB gate columns) for Bc batch rows and keeps its slice of W_hh resident in
// shared memory for all T steps. Per step the CTAs of one (direction, batch-group) exchange the
// small [Bc,H] state through an L2-resident buffer that every CTA pulls into shared memory with a
// 1-D bulk async copy (TMA) tracked by mbarriers, and synchronise with one release/acquire counter.
//
//   fwd step : gates[b, 4UB] = Gx[b,t] + h_{t-1}[b,:] . Wslice^T ; pointwise ; publish h_t slice
//   bwd step : dh = dOut[b,t] + sum_src partial_src[b, my units] ; pointwise -> dG[b,4UB] ;
//              partial_me[b, :] = dG . Wslice  (scattered to every destination's inbox)
//
// The input projection (x . W_ih^T + b_ih + b_hh, K6) and the weight-gradient contractions are plain
// GEMMs done by the caller; this file is the sequential part.
//
// Reference behaviour restated: torch.nn.LSTM as called from /root/reference/src/module.py:112-113,
// 129-132 (single layer, batch_first, zero initial state, run over the zero-padded frames - no
// packing), gate order i,f,g,o.
#include "common.cuh"
#include "../../include/b200asr.h"

namespace b200asr {

constexpr int LSTM_THREADS = 256;
constexpr int LSTM_HALF = 128;
constexpr int LSTM_NCHUNK = 4;

struct LstmParams {
    float* gates;        // [ndir][B][T][H][4]
    const float* whh;    // packed, see lstm_pack kernels
    float* cst;          // [ndir][B][T][H]
    float* out;          // fwd: layer output [B][T][ndir*H]; bwd: dOut (read only)
    float* xbuf;         // exchange buffers
    unsigned* counters;  // [ndir][nbg]
    int* err_flag;
    int B, T, H, ndir, UB, Bc, nub, nbg;
};

__device__ __forceinline__ void spin_until(const unsigned* ctr, unsigned target, int* err_flag) {
    const long long t0 = clock64();
    while (ld_acquire_u32(ctr) < target) {
        if (clock64() - t0 > (1LL << 33)) {  // ~4 s: a peer died; abort instead of hanging the GPU
            *err_flag = 1;
            __threadfence_system();
            __trap();
        }
    }
}

// ------------------------------------------------------------------------------------------
// Weight packing. Source: PyTorch layout w[dir][g*H + j][k] (gate-major rows).
//   fwd pack: dst[dir][ub][k][u][g]           (per-CTA slice contiguous, k-major)
//   bwd pack: dst[dir][ub][u][g][k]           (per-CTA slice contiguous, row-major over k)
__global__ void lstm_pack_kernel(const float* __restrict__ w, float* __restrict__ dst, int H, int UB, int ndir,
                                 int for_bwd) {
    const long long n = (long long)ndir * 4 * H * H;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        // i indexes the destination
        long long r = i;
        const int dir = (int)(r / (4LL * H * H));
        r -= (long long)dir * 4 * H * H;
        const int ub = (int)(r / (4LL * UB * H));
        r -= (long long)ub * 4 * UB * H;
        int k, u, g;
        if (!for_bwd) {
            k = (int)(r / (4 * UB));
            const int q = (int)(r - (long long)k * 4 * UB);
            u = q >> 2;
            g = q & 3;
        } else {
            const int q = (int)(r / H);
            k = (int)(r - (long long)q * H);
            u = q >> 2;
            g = q & 3;
        }
        const int j = ub * UB + u;
        dst[i] = w[((long long)dir * 4 * H + (long long)g * H + j) * H + k];
    }
}

// exchange-buffer index (in float4 units) of (k, bq) inside one [H][Bc] state block:
// chunk-major over k so that each of the LSTM_NCHUNK pieces is one contiguous bulk copy and a
// producer's UB consecutive units land in consecutive 16-B slots.
__device__ __forceinline__ int hx_index(int k, int bq, int KC, int NBQ) {
    const int kc = k / KC;
    return (kc * NBQ + bq) * KC + (k - kc * KC);
}

// ------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(LSTM_THREADS, 1) bilstm_fwd_kernel(LstmParams p) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    const int H = p.H, UB = p.UB, Bc = p.Bc, T = p.T;
    const int NBQ = Bc >> 2;
    const int KC = H / LSTM_NCHUNK;
    float4* Ws = reinterpret_cast<float4*>(s_raw);                           // [H][UB] float4 (4 gates)
    float4* hs = Ws + (size_t)H * UB;                                        // [H*NBQ] float4 (4 batch rows)
    float* red = reinterpret_cast<float*>(hs + (size_t)H * NBQ);             // [16][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + 16 * LSTM_HALF);      // [LSTM_NCHUNK]

    const int tid = threadIdx.x;
    const int half = tid / LSTM_HALF;
    const int pidx = tid % LSTM_HALF;
    int blk = blockIdx.x;
    const int ub = blk % p.nub; blk /= p.nub;
    const int bg = blk % p.nbg; blk /= p.nbg;
    const int dir = blk;

    const int NP = UB * NBQ;
    const bool has_tile = pidx < NP;
    const int u = has_tile ? pidx % UB : 0;
    const int bq = has_tile ? pidx / UB : 0;
    const int ug = ub * UB + u;

    // resident W_hh slice
    {
        const float4* src = reinterpret_cast<const float4*>(p.whh) + ((size_t)dir * p.nub + ub) * (size_t)H * UB;
        for (int i = tid; i < H * UB; i += LSTM_THREADS) Ws[i] = src[i];
    }
    if (tid == 0) {
        for (int c = 0; c < LSTM_NCHUNK; ++c) mbar_init(&bars[c], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const size_t blk_f4 = (size_t)H * NBQ;  // float4 per state block
    float4* xb = reinterpret_cast<float4*>(p.xbuf) + ((size_t)dir * p.nbg + bg) * 2 * blk_f4;
    unsigned* ctr = p.counters + dir * p.nbg + bg;
    const uint32_t chunk_bytes = (uint32_t)((size_t)KC * NBQ * sizeof(float4));

    float c_reg[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t phase = 0;
    const int b0 = bg * Bc + bq * 4;

    for (int step = 0; step < T; ++step) {
        const int tt = dir ? (T - 1 - step) : step;
        float4 gx[4];
        if (half == 0 && has_tile) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int b = b0 + i;
                gx[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (b < p.B)
                    gx[i] = *reinterpret_cast<const float4*>(p.gates + ((((size_t)dir * p.B + b) * T + tt) * H + ug) * 4);
            }
        }
        float acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int g = 0; g < 4; ++g) acc[i][g] = 0.f;

        if (step > 0) {
            if (tid == 0) {
                spin_until(ctr, (unsigned)step * p.nub, p.err_flag);
                fence_proxy_async();
                const float4* src = xb + (size_t)((step - 1) & 1) * blk_f4;
                for (int c = 0; c < LSTM_NCHUNK; ++c) {
                    mbar_expect_tx(&bars[c], chunk_bytes);
                    bulk_g2s(hs + (size_t)c * KC * NBQ, src + (size_t)c * KC * NBQ, chunk_bytes, &bars[c]);
                }
            }
            if (has_tile) {
#pragma unroll 1
                for (int cc = 0; cc < LSTM_NCHUNK / 2; ++cc) {
                    const int c = half * (LSTM_NCHUNK / 2) + cc;
                    mbar_wait(&bars[c], phase);
                    const float4* hp = hs + ((size_t)c * NBQ + bq) * KC;
                    const float4* wp = Ws + (size_t)c * KC * UB + u;
#pragma unroll 4
                    for (int kk = 0; kk < KC; ++kk) {
                        const float4 hv = hp[kk];
                        const float4 wv = wp[(size_t)kk * UB];
                        acc[0][0] = fmaf(hv.x, wv.x, acc[0][0]); acc[0][1] = fmaf(hv.x, wv.y, acc[0][1]);
                        acc[0][2] = fmaf(hv.x, wv.z, acc[0][2]); acc[0][3] = fmaf(hv.x, wv.w, acc[0][3]);
                        acc[1][0] = fmaf(hv.y, wv.x, acc[1][0]); acc[1][1] = fmaf(hv.y, wv.y, acc[1][1]);
                        acc[1][2] = fmaf(hv.y, wv.z, acc[1][2]); acc[1][3] = fmaf(hv.y, wv.w, acc[1][3]);
                        acc[2][0] = fmaf(hv.z, wv.x, acc[2][0]); acc[2][1] = fmaf(hv.z, wv.y, acc[2][1]);
                        acc[2][2] = fmaf(hv.z, wv.z, acc[2][2]); acc[2][3] = fmaf(hv.z, wv.w, acc[2][3]);
                        acc[3][0] = fmaf(hv.w, wv.x, acc[3][0]); acc[3][1] = fmaf(hv.w, wv.y, acc[3][1]);
                        acc[3][2] = fmaf(hv.w, wv.z, acc[3][2]); acc[3][3] = fmaf(hv.w, wv.w, acc[3][3]);
                    }
                }
            } else {
                // threads without a tile still must observe the barriers' phases consistently: nothing to do
            }
            phase ^= 1;
            if (half == 1 && has_tile) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int g = 0; g < 4; ++g) red[(i * 4 + g) * LSTM_HALF + pidx] = acc[i][g];
            }
            __syncthreads();
            if (half == 0 && has_tile) {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int g = 0; g < 4; ++g) acc[i][g] += red[(i * 4 + g) * LSTM_HALF + pidx];
            }
        }

        if (half == 0 && has_tile) {
            float hq[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int b = b0 + i;
                hq[i] = 0.f;
                if (b < p.B) {
                    const float ig = sigmoidf_(gx[i].x + acc[i][0]);
                    const float fg = sigmoidf_(gx[i].y + acc[i][1]);
                    const float gg = tanhf(gx[i].z + acc[i][2]);
                    const float og = sigmoidf_(gx[i].w + acc[i][3]);
                    const float c = fmaf(fg, c_reg[i], ig * gg);
                    c_reg[i] = c;
                    const float h = og * tanhf(c);
                    hq[i] = h;
                    const size_t row = ((size_t)dir * p.B + b) * T + tt;
                    *reinterpret_cast<float4*>(p.gates + (row * H + ug) * 4) = make_float4(ig, fg, gg, og);
                    p.cst[row * H + ug] = c;
                    p.out[((size_t)b * T + tt) * (p.ndir * H) + (size_t)dir * H + ug] = h;
                }
            }
            if (step + 1 < T)
                xb[(size_t)(step & 1) * blk_f4 + hx_index(ug, bq, KC, NBQ)] = make_float4(hq[0], hq[1], hq[2], hq[3]);
        }
        __syncthreads();
        if (tid == 0 && step + 1 < T) {
            __threadfence();
            red_release_add_u32(ctr, 1u);
        }
    }
}

// ------------------------------------------------------------------------------------------
// Backward. Shared memory: Wr[4UB][H] | inbox[nub][Bc][UB] (== Bc*H floats) | dGs[4UB][Bc] | red | bars
__global__ void __launch_bounds__(LSTM_THREADS, 1) bilstm_bwd_kernel(LstmParams p) {
    extern __shared__ __align__(128) unsigned char s_raw[];
    const int H = p.H, UB = p.UB, Bc = p.Bc, T = p.T, nub = p.nub;
    const int NBQ = Bc >> 2;
    float* Wr = reinterpret_cast<float*>(s_raw);                 // [4UB][H]
    float* inbox = Wr + (size_t)4 * UB * H;                      // [nub][Bc][UB]
    float* dGs = inbox + (size_t)Bc * H;                         // [4UB][Bc]
    float* red = dGs + (size_t)4 * UB * Bc;                      // [4][128]
    uint64_t* bars = reinterpret_cast<uint64_t*>(red + 4 * LSTM_HALF);

    const int tid = threadIdx.x;
    const int half = tid / LSTM_HALF;
    const int pidx = tid % LSTM_HALF;
    int blk = blockIdx.x;
    const int ub = blk % nub; blk /= nub;
    const int bg = blk % p.nbg; blk /= p.nbg;
    const int dir = blk;

    const int NP = UB * NBQ;
    const bool has_tile = pidx < NP;
    const int u = has_tile ? pidx % UB : 0;
    const int bq = has_tile ? pidx / UB : 0;
    const int ug = ub * UB + u;
    const int b0 = bg * Bc + bq * 4;

    {
        const float4* src = reinterpret_cast<const float4*>(p.whh) + ((size_t)dir * nub + ub) * (size_t)H * UB;
        float4* dst = reinterpret_cast<float4*>(Wr);
        for (int i = tid; i < H * UB; i += LSTM_THREADS) dst[i] = src[i];
    }
    if (tid == 0) {
        for (int c = 0; c < LSTM_NCHUNK; ++c) mbar_init(&bars[c], 1);
        mbar_fence_init();
    }
    __syncthreads();

    // inbox of destination d (for one parity): [src nub][Bc][UB]  -> Bc*H floats
    const size_t inbox_elems = (size_t)Bc * H;
    float* xb = p.xbuf + ((size_t)dir * p.nbg + bg) * 2 * (size_t)nub * inbox_elems;
    unsigned* ctr = p.counters + dir * p.nbg + bg;
    // the inbox is pulled in LSTM_NCHUNK pieces along the source index
    const int SC = (nub + LSTM_NCHUNK - 1) / LSTM_NCHUNK;  // sources per chunk

    float dc_reg[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t phase = 0;
    const bool vec_ok = (UB % 4) == 0;

    for (int step = 0; step < T; ++step) {
        const int fstep = T - 1 - step;                    // forward step index being differentiated
        const int tt = dir ? (T - 1 - fstep) : fstep;      // its time index
        const int tt_prev = dir ? tt + 1 : tt - 1;         // time index of the previous forward step
        float4 gt[4];
        float ct[4], cp[4], dh[4];
        if (half == 0 && has_tile) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int b = b0 + i;
                gt[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                ct[i] = cp[i] = dh[i] = 0.f;
                if (b < p.B) {
                    const size_t row = ((size_t)dir * p.B + b) * T + tt;
                    gt[i] = *reinterpret_cast<const float4*>(p.gates + (row * H + ug) * 4);
                    ct[i] = p.cst[row * H + ug];
                    if (fstep > 0) cp[i] = p.cst[(((size_t)dir * p.B + b) * T + tt_prev) * H + ug];
                    dh[i] = p.out[((size_t)b * T + tt) * (p.ndir * H) + (size_t)dir * H + ug];
                }
            }
        }
        if (step > 0) {
            if (tid == 0) {
                spin_until(ctr, (unsigned)step * nub, p.err_flag);
                fence_proxy_async();
                const float* src = xb + ((size_t)((step - 1) & 1) * nub + ub) * inbox_elems;
                for (int c = 0; c < LSTM_NCHUNK; ++c) {
                    const int s0 = c * SC;
                    const int s1 = min(nub, s0 + SC);
                    const uint32_t bytes = (s1 > s0) ? (uint32_t)((size_t)(s1 - s0) * Bc * UB * sizeof(float)) : 0u;
                    mbar_expect_tx(&bars[c], bytes);
                    if (bytes) bulk_g2s(inbox + (size_t)s0 * Bc * UB, src + (size_t)s0 * Bc * UB, bytes, &bars[c]);
                }
            }
            float part[4] = {0.f, 0.f, 0.f, 0.f};
            if (has_tile) {
                for (int cc = 0; cc < LSTM_NCHUNK / 2; ++cc) {
                    const int c = half * (LSTM_NCHUNK / 2) + cc;
                    mbar_wait(&bars[c], phase);
                    const int s0 = c * SC;
                    const int s1 = min(nub, s0 + SC);
                    for (int s = s0; s < s1; ++s) {
                        const float* ib = inbox + ((size_t)s * Bc + bq * 4) * UB + u;
#pragma unroll
                        for (int i = 0; i < 4; ++i) part[i] += ib[i * UB];
                    }
                }
            }
            phase ^= 1;
            if (half == 1 && has_tile) {
#pragma unroll
                for (int i = 0; i < 4; ++i) red[i * LSTM_HALF + pidx] = part[i];
            }
            __syncthreads();
            if (half == 0 && has_tile) {
#pragma unroll
                for (int i = 0; i < 4; ++i) dh[i] += part[i] + red[i * LSTM_HALF + pidx];
            }
        }
        // pointwise backward of the cell
        if (half == 0 && has_tile) {
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int b = b0 + i;
                float4 dg = make_float4(0.f, 0.f, 0.f, 0.f);
                if (b < p.B) {
                    const float ig = gt[i].x, fg = gt[i].y, gg = gt[i].z, og = gt[i].w;
                    const float tc = tanhf(ct[i]);
                    const float dout_o = dh[i] * tc;
                    const float dc = dc_reg[i] + dh[i] * og * (1.f - tc * tc);
                    dg.x = dc * gg * ig * (1.f - ig);
                    dg.y = dc * cp[i] * fg * (1.f - fg);
                    dg.z = dc * ig * (1.f - gg * gg);
                    dg.w = dout_o * og * (1.f - og);
                    dc_reg[i] = dc * fg;
                    const size_t row = ((size_t)dir * p.B + b) * T + tt;
                    *reinterpret_cast<float4*>(p.gates + (row * H + ug) * 4) = dg;
                }
                // stage for the matmul: dGs[(u*4+g)][b_local]
                const int bl = bq * 4 + i;
                dGs[(u * 4 + 0) * Bc + bl] = dg.x;
                dGs[(u * 4 + 1) * Bc + bl] = dg.y;
                dGs[(u * 4 + 2) * Bc + bl] = dg.z;
                dGs[(u * 4 + 3) * Bc + bl] = dg.w;
            }
        }
        __syncthreads();
        if (step + 1 < T) {
            // partial[b][k] = sum_c dGs[c][b] * Wr[c][k]; tiles of 4 batch rows x (4 strided float4 of k)
            const int NKQ = H / 16;              // threads along k; each owns k = r*(H/4) + kq*4 + j
            const int ntiles = NKQ * NBQ;
            float* outbase = xb + (size_t)(step & 1) * nub * inbox_elems;
            for (int tile = tid; tile < ntiles; tile += LSTM_THREADS) {
                const int kq = tile % NKQ;
                const int tbq = tile / NKQ;
                float a[4][16];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 16; ++j) a[i][j] = 0.f;
                const int C = 4 * UB;
#pragma unroll 2
                for (int c = 0; c < C; ++c) {
                    const float4 d4 = *reinterpret_cast<const float4*>(dGs + (size_t)c * Bc + tbq * 4);
                    const float dv[4] = {d4.x, d4.y, d4.z, d4.w};
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const float4 w4 = *reinterpret_cast<const float4*>(Wr + (size_t)c * H + r * (H / 4) + kq * 4);
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            a[i][r * 4 + 0] = fmaf(dv[i], w4.x, a[i][r * 4 + 0]);
                            a[i][r * 4 + 1] = fmaf(dv[i], w4.y, a[i][r * 4 + 1]);
                            a[i][r * 4 + 2] = fmaf(dv[i], w4.z, a[i][r * 4 + 2]);
                            a[i][r * 4 + 3] = fmaf(dv[i], w4.w, a[i][r * 4 + 3]);
                        }
                    }
                }
                // scatter to the destination inboxes: element (dst, src=ub, b_local, u')
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const int k0 = r * (H / 4) + kq * 4;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int bl = tbq * 4 + i;
                        if (vec_ok) {
                            const int dst = k0 / UB, uu = k0 - dst * UB;
                            float* o = outbase + (((size_t)dst * nub + ub) * Bc + bl) * UB + uu;
                            *reinterpret_cast<float4*>(o) =
                                make_float4(a[i][r * 4], a[i][r * 4 + 1], a[i][r * 4 + 2], a[i][r * 4 + 3]);
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int k = k0 + j;
                                const int dst = k / UB, uu = k - dst * UB;
                                outbase[(((size_t)dst * nub + ub) * Bc + bl) * UB + uu] = a[i][r * 4 + j];
                            }
                        }
                    }
                }
            }
            __syncthreads();
            if (tid == 0) {
                __threadfence();
                red_release_add_u32(ctr, 1u);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------
__global__ void lstm_cell_fwd_kernel(const float* __restrict__ pre, const float* __restrict__ c_prev,
                                     float* __restrict__ gates, float* __restrict__ c, float* __restrict__ h, int B,
                                     int H) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * H) return;
    const int b = i / H, j = i - b * H;
    const float* pr = pre + (size_t)b * 4 * H;
    const float ig = sigmoidf_(pr[j]), fg = sigmoidf_(pr[H + j]), gg = tanhf(pr[2 * H + j]),
                og = sigmoidf_(pr[3 * H + j]);
    const float cn = fmaf(fg, c_prev[i], ig * gg);
    float* gr = gates + (size_t)b * 4 * H;
    gr[j] = ig; gr[H + j] = fg; gr[2 * H + j] = gg; gr[3 * H + j] = og;
    c[i] = cn;
    h[i] = og * tanhf(cn);
}

__global__ void lstm_cell_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev,
                                     const float* __restrict__ c, const float* __restrict__ dh,
                                     const float* __restrict__ dc_next, float* __restrict__ dpre,
                                     float* __restrict__ dc_prev, int B, int H) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B * H) return;
    const int b = i / H, j = i - b * H;
    const float* gr = gates + (size_t)b * 4 * H;
    const float ig = gr[j], fg = gr[H + j], gg = gr[2 * H + j], og = gr[3 * H + j];
    const float tc = tanhf(c[i]);
    const float dhv = dh[i];
    const float dc = (dc_next ? dc_next[i] : 0.f) + dhv * og * (1.f - tc * tc);
    float* dp = dpre + (size_t)b * 4 * H;
    dp[j] = dc * gg * ig * (1.f - ig);
    dp[H + j] = dc * c_prev[i] * fg * (1.f - fg);
    dp[2 * H + j] = dc * ig * (1.f - gg * gg);
    dp[3 * H + j] = dhv * tc * og * (1.f - og);
    dc_prev[i] = dc * fg;
}

// ------------------------------------------------------------------------------------------
struct Plan {
    int UB, Bc, nub, nbg, ctas;
    size_t smem_fwd, smem_bwd, pack_bytes, xbuf_fwd_bytes, xbuf_bwd_bytes;
};

static size_t smem_fwd_bytes(int H, int UB, int Bc) {
    return (size_t)H * UB * 16 + (size_t)H * (Bc / 4) * 16 + 16 * LSTM_HALF * 4 + LSTM_NCHUNK * 8 + 128;
}
static size_t smem_bwd_bytes(int H, int UB, int Bc) {
    return (size_t)4 * UB * H * 4 + (size_t)Bc * H * 4 + (size_t)4 * UB * Bc * 4 + 4 * LSTM_HALF * 4 +
           LSTM_NCHUNK * 8 + 128;
}

static int make_plan(int B, int H, int ndir, Plan* out) {
    if (H % 16 != 0) return -1;  // K chunks (H/4) and the 16-wide k tiles of the backward pass
    const int sms = sm_count();
    const size_t smem_cap = (size_t)max_optin_smem();
    long long best_cost = -1;
    Plan best{};
    for (int UB = 1; UB <= H; ++UB) {
        if (H % UB) continue;
        for (int Bc = 4; Bc <= 64; Bc += 4) {
            const int NP = UB * (Bc / 4);
            if (NP > LSTM_HALF) continue;
            const int nbg = (B + Bc - 1) / Bc;
            const int nub = H / UB;
            const int ctas = ndir * nbg * nub;
            if (ctas > sms) continue;
            const size_t sf = smem_fwd_bytes(H, UB, Bc), sb = smem_bwd_bytes(H, UB, Bc);
            if (sf > smem_cap || sb > smem_cap) continue;
            // bulk copies need 16-B multiples
            if (((size_t)(H / LSTM_NCHUNK) * (Bc / 4) * 16) % 16) continue;
            if (((size_t)Bc * UB * 4) % 16) continue;
            // cost: per-CTA FMA work per step; tie-break on the state tile pulled per step
            const long long cost = (long long)UB * Bc * 1000 + Bc;
            if (best_cost < 0 || cost < best_cost) {
                best_cost = cost;
                best.UB = UB; best.Bc = Bc; best.nub = nub; best.nbg = nbg; best.ctas = ctas;
                best.smem_fwd = sf; best.smem_bwd = sb;
            }
        }
    }
    if (best_cost < 0) return -2;
    best.pack_bytes = (size_t)ndir * 4 * H * H * sizeof(float);
    best.xbuf_fwd_bytes = (size_t)ndir * best.nbg * 2 * (size_t)H * best.Bc * sizeof(float);
    best.xbuf_bwd_bytes = (size_t)ndir * best.nbg * 2 * (size_t)best.nub * best.Bc * H * sizeof(float);
    *out = best;
    return 0;
}

static size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

}  // namespace b200asr

using namespace b200asr;

extern "C" size_t b200asr_bilstm_workspace_bytes(int B, int T, int H, int ndir) {
    (void)T;
    Plan pl;
    if (make_plan(B, H, ndir, &pl) != 0) return 0;
    const size_t x = pl.xbuf_fwd_bytes > pl.xbuf_bwd_bytes ? pl.xbuf_fwd_bytes : pl.xbuf_bwd_bytes;
    return align_up(pl.pack_bytes, 256) + align_up(x, 256) + 256 /*counters + err flag*/;
}

extern "C" int b200asr_bilstm_plan(int B, int H, int ndir, int* unit_block, int* batch_block, int* n_ctas) {
    Plan pl;
    const int rc = make_plan(B, H, ndir, &pl);
    B200_REQUIRE(rc == 0, "bilstm_plan: no feasible decomposition for B=%d H=%d ndir=%d (H must be a multiple of 16)",
                 B, H, ndir);
    if (unit_block) *unit_block = pl.UB;
    if (batch_block) *batch_block = pl.Bc;
    if (n_ctas) *n_ctas = pl.ctas;
    return B200_OK;
}

static int bilstm_run(bool bwd, float* gates, const float* w_hh, float* cstate, float* out_or_dout, int B, int T,
                      int H, int ndir, void* workspace, size_t workspace_bytes, cudaStream_t stream) {
    B200_REQUIRE(gates && w_hh && cstate && out_or_dout && workspace, "bilstm: null pointer");
    B200_REQUIRE(B > 0 && T > 0 && H > 0 && (ndir == 1 || ndir == 2), "bilstm: bad sizes B=%d T=%d H=%d ndir=%d", B, T,
                 H, ndir);
    Plan pl;
    B200_REQUIRE(make_plan(B, H, ndir, &pl) == 0,
                 "bilstm: no feasible decomposition for B=%d H=%d ndir=%d (H must be a multiple of 16)", B, H, ndir);
    B200_REQUIRE(workspace_bytes >= b200asr_bilstm_workspace_bytes(B, T, H, ndir), "bilstm: workspace too small");
    unsigned char* ws = reinterpret_cast<unsigned char*>(workspace);
    float* packed = reinterpret_cast<float*>(ws);
    const size_t xoff = align_up(pl.pack_bytes, 256);
    const size_t xbytes = pl.xbuf_fwd_bytes > pl.xbuf_bwd_bytes ? pl.xbuf_fwd_bytes : pl.xbuf_bwd_bytes;
    float* xbuf = reinterpret_cast<float*>(ws + xoff);
    unsigned* counters = reinterpret_cast<unsigned*>(ws + xoff + align_up(xbytes, 256));
    int* err_flag = reinterpret_cast<int*>(counters + 32);

    B200_CUDA(cudaMemsetAsync(counters, 0, 256, stream));
    {
        const long long n = (long long)ndir * 4 * H * H;
        int blocks = (int)((n + 255) / 256);
        if (blocks > 4096) blocks = 4096;
        lstm_pack_kernel<<<blocks, 256, 0, stream>>>(w_hh, packed, H, pl.UB, ndir, bwd ? 1 : 0);
        B200_LAUNCH_CHECK("lstm_pack_kernel");
    }
    LstmParams p;
    p.gates = gates; p.whh = packed; p.cst = cstate; p.out = out_or_dout; p.xbuf = xbuf; p.counters = counters;
    p.err_flag = err_flag; p.B = B; p.T = T; p.H = H; p.ndir = ndir; p.UB = pl.UB; p.Bc = pl.Bc; p.nub = pl.nub;
    p.nbg = pl.nbg;
    const void* fn = bwd ? (const void*)bilstm_bwd_kernel : (const void*)bilstm_fwd_kernel;
    const size_t smem = bwd ? pl.smem_bwd : pl.smem_fwd;
    B200_CUDA(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    B200_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, fn, LSTM_THREADS, smem));
    B200_REQUIRE((long long)per_sm * sm_count() >= pl.ctas, "bilstm: %d CTAs cannot be co-resident (%d/SM x %d SMs)",
                 pl.ctas, per_sm, sm_count());
    void* args[] = {&p};
    B200_CUDA(cudaLaunchCooperativeKernel(fn, dim3(pl.ctas), dim3(LSTM_THREADS), args, smem, stream));
    count_launch();
    return B200_OK;
}

extern "C" int b200asr_bilstm_fwd(float* gates, const float* w_hh, float* cstate, float* out, int B, int T, int H,
                                  int ndir, void* workspace, size_t workspace_bytes, b200asr_stream stream) {
    return bilstm_run(false, gates, w_hh, cstate, out, B, T, H, ndir, workspace, workspace_bytes,
                      (cudaStream_t)stream);
}

extern "C" int b200asr_bilstm_bwd(float* gates, const float* w_hh, const float* cstate, const float* dout, int B,
                                  int T, int H, int ndir, void* workspace, size_t workspace_bytes,
                                  b200asr_stream stream) {
    return bilstm_run(true, gates, w_hh, const_cast<float*>(cstate), const_cast<float*>(dout), B, T, H, ndir,
                      workspace, workspace_bytes, (cudaStream_t)stream);
}

extern "C" int b200asr_lstm_cell_fwd(const float* preact, const float* c_prev, float* gates, float* c, float* h,
                                     int B, int H, b200asr_stream stream) {
    B200_REQUIRE(preact && c_prev && gates && c && h, "lstm_cell_fwd: null pointer");
    B200_REQUIRE(B > 0 && H > 0, "lstm_cell_fwd: bad sizes");
    const int n = B * H;
    lstm_cell_fwd_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(preact, c_prev, gates, c, h, B, H);
    B200_LAUNCH_CHECK("lstm_cell_fwd_kernel");
    return B200_OK;
}

extern "C" int b200asr_lstm_cell_bwd(const float* gates, const float* c_prev, const float* c, const float* dh,
                                     const float* dc_next, float* dpreact, float* dc_prev, int B, int H,
                                     b200asr_stream stream) {
    B200_REQUIRE(gates && c_prev && c && dh && dpreact && dc_prev, "lstm_cell_bwd: null pointer");
    B200_REQUIRE(B > 0 && H > 0, "lstm_cell_bwd: bad sizes");
    const int n = B * H;
    lstm_cell_bwd_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(gates, c_prev, c, dh, dc_next, dpreact,
                                                                          dc_prev, B, H);
    B200_LAUNCH_CHECK("lstm_cell_bwd_kernel");
    return B200_OK;
}
