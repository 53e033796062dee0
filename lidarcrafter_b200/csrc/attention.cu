// K2: attention cores of the two denoisers, softmax(q k^T * scale) v per (batch, head), flash style (online softmax; the
// scores never leave the SM, unlike the reference's [B*h, T, S] fp32 tensor).
//   MHA : q/k/v = column slices of the fused qkv tensor (nn.MultiheadAttention core, efficient_unet.py:39-53), d = dv = 32 | 64
//   OA  : q = [q_c ; pos_p], k = [[k_c ; pos_p] | [k_l ; pos_l]], v = [v_c | v_l], d = 32 + 32, dv = 32
//         (ObjectAwareCrossAttention / QKVAttentionLegacy, layout_unet_v1.py:416-532)
// Product path: flash_attn_tc_kernel (tcgen05.mma, accumulators in TMEM, operands staged by the TMA engine) behind
// attn_pack_kernel; flash_attn_mma_kernel (mma.sync, register-level) is the cross-check / A-B implementation
// (B200_FA_IMPL=mma).  Output: fp16 hi | lo (or e4m3 pair) conv operand of the out-projection.
#include <stdlib.h>

#include <cuda_bf16.h>

#include "common.cuh"

namespace b200 {

struct FAParams {
    // q/k/v "source 1" (content) and optional "source 2" (positional) pointers; token-major fp32
    const float *q1, *q2, *k1, *k2, *v;      // main tokens
    const float *xk1, *xk2, *xv;             // extra (layout) keys / values, may be null
    int ldq1, ldq2, ldk1, ldk2, ldv, ldx;    // row strides (floats)
    int d1, d2, dv;                          // per-head dims of source 1 / 2 and of v
    int T, Tx;                               // #queries == #main keys, #extra keys
    __half* out;
    size_t lo_off;   // fp16 elements per operand plane
    int parts;
    int C, Wimg;                             // output channels (heads * dv) and image width for the operand store
    float scale;                             // applied to q (full softmax scale)
    __half* ws;                              // tcgen05 path: packed Q / K / V^T tile images (attn_pack_kernel), else unused
};

// ---------------------------------------------------------------------------------------------------------
// warp = 16 query rows.  S = Q K^T and O += P V run on mma.sync.m16n8k16 (fp16 operands, fp32 accumulators in
// registers) with the error-compensated split used by the convs: x = x_hi + x_lo (both fp16), three MMAs per product
// (hi*hi + lo*hi + hi*lo), so logits and outputs keep ~2^-22 relative accuracy -- a single fp16 pass would put 1e-2
// absolute error on logits of magnitude ~10.  P never leaves registers: the accumulator fragment of two adjacent
// 8-key tiles IS the A fragment of the next k-step of P V.  K is staged [key][d] and V transposed [dv][key] (both
// hi | lo planes, rows padded by 8 halves -> conflict-free 32-bit fragment loads).  Softmax runs in the exp2 domain
// (scale * log2 e folded into Q before the split).  (A tcgen05/TMEM version would need the FA4 correction pipeline;
// at T <= 2048 and d <= 64 the step is 6-15% attention, so the register-level path was chosen.)
// ---------------------------------------------------------------------------------------------------------
constexpr int FM_BQ = 64, FM_BK = 64, FM_THREADS = 128;

__device__ __forceinline__ void mma_f16_16816(float* d, const uint32_t* a, uint32_t b0, uint32_t b1) {
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// (x, y) -> fp16 pair hi and the fp16 pair of the residuals
__device__ __forceinline__ void split_h2(float x, float y, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x, y);
    const float2 hf = __half22float2(h);
    const __half2 l = __floats2half2_rn(x - hf.x, y - hf.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}

template <int DQ, int DV>
__global__ void __launch_bounds__(FM_THREADS) flash_attn_mma_kernel(const FAParams p) {
    constexpr int KP = DQ + 8, VP = FM_BK + 8;
    extern __shared__ __align__(16) unsigned char fm_smem[];
    pdl_launch_dependents();
    pdl_wait();
    __half* sKh = reinterpret_cast<__half*>(fm_smem);   // [FM_BK][KP]
    __half* sKl = sKh + FM_BK * KP;
    __half* sVh = sKl + FM_BK * KP;                      // [DV][VP]  (transposed: value dim major, key minor)
    __half* sVl = sVh + DV * VP;
    const int q0 = blockIdx.x * FM_BQ, head = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const float qscale = p.scale * 1.4426950408889634f;

    // ---- this warp's 16 query rows as A fragments (hi | lo), scaled ----
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;
    auto qpair = [&](int row, int c, uint32_t& hi, uint32_t& lo) {   // elements (row, c) and (row, c + 1)
        float2 v = make_float2(0.f, 0.f);
        if (row < p.T) {
            const size_t tok = (size_t)b * p.T + row;
            v = c < p.d1 ? *reinterpret_cast<const float2*>(p.q1 + tok * p.ldq1 + head * p.d1 + c)
                         : *reinterpret_cast<const float2*>(p.q2 + tok * p.ldq2 + head * p.d2 + (c - p.d1));
        }
        split_h2(v.x * qscale, v.y * qscale, hi, lo);
    };
    uint32_t qh[DQ / 16][4], ql[DQ / 16][4];
#pragma unroll
    for (int ks = 0; ks < DQ / 16; ++ks) {
        const int c = ks * 16 + 2 * t;
        qpair(r0, c, qh[ks][0], ql[ks][0]);
        qpair(r1, c, qh[ks][1], ql[ks][1]);
        qpair(r0, c + 8, qh[ks][2], ql[ks][2]);
        qpair(r1, c + 8, qh[ks][3], ql[ks][3]);
    }
    float o[DV / 8][4];
#pragma unroll
    for (int n = 0; n < DV / 8; ++n) o[n][0] = o[n][1] = o[n][2] = o[n][3] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;    // rows g and g + 8 (l: this thread's partial sums)

    const int S = p.T + p.Tx;
    const int ntiles = (S + FM_BK - 1) / FM_BK;
    // K / V tiles travel global -> registers (requested one tile ahead, in flight during the MMAs of the current tile)
    // -> fp16 hi | lo split -> shared memory
    constexpr int KN = FM_BK * (DQ / 4) / FM_THREADS, VN = FM_BK * (DV / 4) / FM_THREADS;
    float4 kreg[KN], vreg[VN];
    auto load_tile = [&](int k0) {
#pragma unroll
        for (int u = 0; u < KN; ++u) {
            const int i = tid + u * FM_THREADS;
            const int kj = i / (DQ / 4), c4 = (i - kj * (DQ / 4)) * 4;
            const int tk = k0 + kj;
            float4 v = make_float4(0, 0, 0, 0);
            if (tk < p.T) {
                const size_t tok = (size_t)b * p.T + tk;
                v = c4 < p.d1 ? *reinterpret_cast<const float4*>(p.k1 + tok * p.ldk1 + head * p.d1 + c4)
                              : *reinterpret_cast<const float4*>(p.k2 + tok * p.ldk2 + head * p.d2 + (c4 - p.d1));
            } else if (tk < S) {
                const size_t tok = (size_t)b * p.Tx + (tk - p.T);
                v = c4 < p.d1 ? *reinterpret_cast<const float4*>(p.xk1 + tok * p.ldx + head * p.d1 + c4)
                              : *reinterpret_cast<const float4*>(p.xk2 + tok * p.ldx + head * p.d2 + (c4 - p.d1));
            }
            kreg[u] = v;
        }
#pragma unroll
        for (int u = 0; u < VN; ++u) {
            const int i = tid + u * FM_THREADS;
            const int kj = i % FM_BK, c4 = (i / FM_BK) * 4;            // consecutive threads -> consecutive keys
            const int tk = k0 + kj;
            float4 v = make_float4(0, 0, 0, 0);
            if (tk < p.T) v = *reinterpret_cast<const float4*>(p.v + ((size_t)b * p.T + tk) * p.ldv + head * DV + c4);
            else if (tk < S) v = *reinterpret_cast<const float4*>(p.xv + ((size_t)b * p.Tx + (tk - p.T)) * p.ldx + head * DV + c4);
            vreg[u] = v;
        }
    };
    load_tile(0);
    for (int kt = 0; kt < ntiles; ++kt) {
        const int k0 = kt * FM_BK;
        __syncthreads();       // previous tile fully consumed
        // ---- registers -> K [key][d] and V^T [dv][key], split into fp16 hi | lo ----
#pragma unroll
        for (int u = 0; u < KN; ++u) {
            const int i = tid + u * FM_THREADS;
            const int kj = i / (DQ / 4), c4 = (i - kj * (DQ / 4)) * 4;
            uint2 hi, lo;
            split_h2(kreg[u].x, kreg[u].y, hi.x, lo.x);
            split_h2(kreg[u].z, kreg[u].w, hi.y, lo.y);
            *reinterpret_cast<uint2*>(sKh + kj * KP + c4) = hi;
            *reinterpret_cast<uint2*>(sKl + kj * KP + c4) = lo;
        }
#pragma unroll
        for (int u = 0; u < VN; ++u) {
            const int i = tid + u * FM_THREADS;
            const int kj = i % FM_BK, c4 = (i / FM_BK) * 4;
            const float vv[4] = {vreg[u].x, vreg[u].y, vreg[u].z, vreg[u].w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const __half hi = __float2half_rn(vv[e]);
                sVh[(c4 + e) * VP + kj] = hi;
                sVl[(c4 + e) * VP + kj] = __float2half_rn(vv[e] - __half2float(hi));
            }
        }
        __syncthreads();
        if (kt + 1 < ntiles) load_tile(k0 + FM_BK);
        // ---- S = Q K^T: 8 key tiles of 8, fp32 accumulators ----
        // (loop order: consecutive MMAs always target DIFFERENT accumulators -- the three split terms of one accumulator
        //  are 4 instructions apart -- because back-to-back dependent HMMAs stall for the full pipeline latency)
        float s[8][4];
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j][0] = s[j][1] = s[j][2] = s[j][3] = 0.f;
#pragma unroll
        for (int ks = 0; ks < DQ / 16; ++ks) {
#pragma unroll
            for (int j0 = 0; j0 < 8; j0 += 4) {
                uint32_t bh[4][2], bl[4][2];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const __half* kh = sKh + ((j0 + jj) * 8 + g) * KP + 2 * t + ks * 16;
                    const __half* kl = sKl + ((j0 + jj) * 8 + g) * KP + 2 * t + ks * 16;
                    bh[jj][0] = *reinterpret_cast<const uint32_t*>(kh);
                    bh[jj][1] = *reinterpret_cast<const uint32_t*>(kh + 8);
                    bl[jj][0] = *reinterpret_cast<const uint32_t*>(kl);
                    bl[jj][1] = *reinterpret_cast<const uint32_t*>(kl + 8);
                }
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) mma_f16_16816(s[j0 + jj], qh[ks], bh[jj][0], bh[jj][1]);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) mma_f16_16816(s[j0 + jj], ql[ks], bh[jj][0], bh[jj][1]);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) mma_f16_16816(s[j0 + jj], qh[ks], bl[jj][0], bl[jj][1]);
            }
        }
        if (k0 + FM_BK > S) {      // mask keys past the end (last tile only)
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int kc = k0 + j * 8 + 2 * t;
                if (kc >= S) s[j][0] = s[j][2] = -INFINITY;
                if (kc + 1 >= S) s[j][1] = s[j][3] = -INFINITY;
            }
        }
        // ---- online softmax (exp2 domain); a row lives in the 4 lanes of a quad ----
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            mx0 = fmaxf(mx0, fmaxf(s[j][0], s[j][1]));
            mx1 = fmaxf(mx1, fmaxf(s[j][2], s[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1)); mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1)); mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float al0 = exp2f(m0 - mn0), al1 = exp2f(m1 - mn1);     // exp2(-inf) = 0 on the first tile
        m0 = mn0; m1 = mn1;
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            s[j][0] = exp2f(s[j][0] - mn0); s[j][1] = exp2f(s[j][1] - mn0);
            s[j][2] = exp2f(s[j][2] - mn1); s[j][3] = exp2f(s[j][3] - mn1);
            rs0 += s[j][0] + s[j][1];
            rs1 += s[j][2] + s[j][3];
        }
        l0 = l0 * al0 + rs0;
        l1 = l1 * al1 + rs1;
#pragma unroll
        for (int n = 0; n < DV / 8; ++n) { o[n][0] *= al0; o[n][1] *= al0; o[n][2] *= al1; o[n][3] *= al1; }
        // ---- O += P V: the S fragments of key tiles (2kk, 2kk+1) are the A fragment of k-step kk ----
#pragma unroll
        for (int kk = 0; kk < FM_BK / 16; ++kk) {
            uint32_t ph[4], pl[4];
            split_h2(s[2 * kk][0], s[2 * kk][1], ph[0], pl[0]);
            split_h2(s[2 * kk][2], s[2 * kk][3], ph[1], pl[1]);
            split_h2(s[2 * kk + 1][0], s[2 * kk + 1][1], ph[2], pl[2]);
            split_h2(s[2 * kk + 1][2], s[2 * kk + 1][3], ph[3], pl[3]);
#pragma unroll
            for (int n0 = 0; n0 < DV / 8; n0 += 4) {
                uint32_t bh[4][2], bl[4][2];
#pragma unroll
                for (int nn = 0; nn < 4; ++nn) {
                    const __half* vh = sVh + ((n0 + nn) * 8 + g) * VP + kk * 16 + 2 * t;
                    const __half* vl = sVl + ((n0 + nn) * 8 + g) * VP + kk * 16 + 2 * t;
                    bh[nn][0] = *reinterpret_cast<const uint32_t*>(vh);
                    bh[nn][1] = *reinterpret_cast<const uint32_t*>(vh + 8);
                    bl[nn][0] = *reinterpret_cast<const uint32_t*>(vl);
                    bl[nn][1] = *reinterpret_cast<const uint32_t*>(vl + 8);
                }
#pragma unroll
                for (int nn = 0; nn < 4; ++nn) mma_f16_16816(o[n0 + nn], ph, bh[nn][0], bh[nn][1]);
#pragma unroll
                for (int nn = 0; nn < 4; ++nn) mma_f16_16816(o[n0 + nn], pl, bh[nn][0], bh[nn][1]);
#pragma unroll
                for (int nn = 0; nn < 4; ++nn) mma_f16_16816(o[n0 + nn], ph, bl[nn][0], bl[nn][1]);
            }
        }
    }
    // ---- normalise and store (conv operand layout, common.cuh) ----
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1); l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1); l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
#pragma unroll
    for (int hrow = 0; hrow < 2; ++hrow) {
        const int tq = hrow ? r1 : r0;
        if (tq >= p.T) continue;
        const float inv = hrow ? inv1 : inv0;
        const int hh = tq / p.Wimg, ww = tq - hh * p.Wimg;
#pragma unroll
        for (int n = 0; n < DV / 8; ++n)
#pragma unroll
            for (int e = 0; e < 2; ++e)
                store_operand_elem(p.out, p.lo_off, p.parts, (size_t)b * (p.T / p.Wimg) + hh, p.C, p.Wimg, ww,
                                   head * DV + n * 8 + 2 * t + e, o[n][2 * hrow + e] * inv);
    }
}


// =========================================================================================================
// tcgen05 / TMEM flash attention (the product path): one CTA = 128 queries of one (batch, head); key blocks of 128.
//   S = Q K^T  : tcgen05.mma kind::f16, M = 128 (queries) x N = 128 (keys) x K = DQ, fp32 accumulator in TMEM (two S
//                buffers: the MMAs of block j + 1 run while the softmax warps work on block j);
//   softmax    : warps 0-3, thread i owns query row i = TMEM lane i: tcgen05.ld of the row (two passes: running max, then
//                exp2 / row sum), P split into fp16 hi | lo and written to shared memory as the K-major A operand of
//   O_j = P V  : tcgen05.mma M = 128 x N = 2 DV x K = 128 keys into a second TMEM region; the rows of the staged V^T tile
//                are [V_hi ; V_lo], so P_hi x [V_hi ; V_lo] is ONE MMA per k-step (+ P_lo x V_hi with N = DV): the softmax
//                warps fold O_j into their fp32 register accumulator with the online-softmax rescale (no TMEM
//                read-modify-write, no correction warpgroup);
//   loaders    : warps 4-7 convert the fp32 K / V rows (and the layout keys / values of the object-aware variant) to
//                fp16 hi | lo straight into the canonical no-swizzle K-major shared-memory image (two stages each);
//   issuer     : warp 8, one elected lane.
// Same error-compensated arithmetic as flash_attn_mma_kernel (three fp16 products per term, ~2^-22 relative): logits of
// magnitude ~10 would carry 1e-2 absolute error in a single fp16 pass.  mbarrier protocol: K/V FULL (128 loader arrivals)
// / EMPTY (tcgen05.commit), S FULL (commit) / EMPTY (128 softmax arrivals), P FULL (128), O FULL (commit; also frees P and
// the V stage) / EMPTY (128).
// =========================================================================================================
constexpr int FT_BQ = 128, FT_BK = 128;
constexpr int FT_NP = 4;                 // softmax threads per query row (each owns 128 / FT_NP key columns of a block)
constexpr int FT_SOFTMAX = 128 * FT_NP;  // warps 0-15: warps w, w + 4, w + 8, w + 12 share the query rows of TMEM lane quarter w % 4
constexpr int FT_THREADS = FT_SOFTMAX + 64;   // + TMA producer warp, MMA issuer warp
constexpr int FT_WP = FT_SOFTMAX / 32, FT_WM = FT_WP + 1;     // their warp indices

// optional in-kernel profile (b200_attn_set_debug): 16 x u64 cycle counters per CTA
//   [0] issuer total [1] wait K [2] wait S free [3] wait P [4] wait V [5] wait O free
//   [6] softmax thread 0 total [7] wait S [8] wait O (fold) [9] row-max exchange barrier
//   [10] loader thread 0 total [11] wait K stage free [12] wait V stage free [13] K load + convert [14] V load + convert
__device__ unsigned long long* g_attn_dbg = nullptr;
#define FT_T0() const long long t0__ = dbg ? clock64() : 0
#define FT_ACC(slot) do { if (dbg) dbg_acc[slot] += clock64() - t0__; } while (0)

template <int DQ, int DV>
struct FTCfg {
    static constexpr int QSLAB = FT_BQ * 16;                 // one 8-column group of the Q tile (= LBO of its descriptor)
    static constexpr int Q_PLANE = (DQ / 8) * QSLAB, Q_BYTES = 2 * Q_PLANE;
    static constexpr int KSLAB = FT_BK * 16;
    static constexpr int K_PLANE = (DQ / 8) * KSLAB, K_STAGE = 2 * K_PLANE;
    static constexpr int VROWS = 2 * DV;                     // rows [V_hi (DV) ; V_lo (DV)] of the transposed value tile
    static constexpr int VSLAB = VROWS * 16;                 // one group of 8 keys
    static constexpr int V_STAGE = (FT_BK / 8) * VSLAB;
    static constexpr int PSLAB = FT_BQ * 16;
    static constexpr int P_PLANE = (FT_BK / 8) * PSLAB, P_BYTES = 2 * P_PLANE;
    static constexpr int OFF_Q = 0, OFF_K = Q_BYTES, OFF_V = OFF_K + 2 * K_STAGE, OFF_P = OFF_V + 2 * V_STAGE;
    static constexpr int OFF_X = OFF_P + P_BYTES;            // row-maximum / row-sum exchange of the two column halves
    static constexpr int OFF_BAR = OFF_X + FT_NP * FT_BQ * 4;      // bf16 [2 parities][FT_NP][128] maxima, reused as fp32 [FT_NP][128] sums
    static constexpr int SMEM = OFF_BAR + 192;
    static constexpr int S_COLS = FT_BK, O_COLS = 2 * DV;    // TMEM: S buffers at columns 0 / 128, O buffers at 256 / 256 + 2 DV
    static constexpr int TMEM_COLS = 512;
    static_assert(SMEM <= 227 * 1024 && 2 * S_COLS + 2 * O_COLS <= TMEM_COLS, "flash_attn_tc budget");
};

__device__ __forceinline__ float ex2_f(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}
// 8 fp32 values -> one 16-byte unit of fp16 hi and one of fp16 lo
__device__ __forceinline__ void split8_store(const float* v, uint32_t a_hi, uint32_t a_lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_h2(v[2 * e], v[2 * e + 1], h[e], l[e]);
    sts_v4(a_hi, h[0], h[1], h[2], h[3]);
    sts_v4(a_lo, l[0], l[1], l[2], l[3]);
}
// 32 TMEM columns of this thread's lane, without the wait (issue several, then tmem_ld_wait())
__device__ __forceinline__ void tmem_ld_32x32_nowait(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x16_nowait(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_32x8_nowait(uint32_t taddr, float* v) {
    uint32_t* r = reinterpret_cast<uint32_t*>(v);
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// 8 consecutive fp32 of a query / key row (two 16-byte loads), or zeros
__device__ __forceinline__ void load8(const float* src, float* v, float scale) {
    if (src) {
        const float4 a = *reinterpret_cast<const float4*>(src), b = *reinterpret_cast<const float4*>(src + 4);
        v[0] = a.x * scale; v[1] = a.y * scale; v[2] = a.z * scale; v[3] = a.w * scale;
        v[4] = b.x * scale; v[5] = b.y * scale; v[6] = b.z * scale; v[7] = b.w * scale;
    } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = 0.f;
    }
}

// ---------------------------------------------------------------------------------------------------------
// pack pass of the tcgen05 attention: fp32 q / k / v rows (content | positional parts, image | layout tokens) -> fp16 hi | lo
// tile images in exactly the shared-memory layout the MMAs read (K-major, no swizzle):
//   Q image of (b, head, q-tile)   [plane][DQ/8][128 rows][8]            pre-scaled by softmax scale * log2 e
//   K image of (b, head, block j)  [plane][DQ/8][128 keys][8]
//   V image of (b, head, block j)  [16 key groups][V_hi rows (DV) ; V_lo rows (DV)][8 keys]   (transposed)
// Keys / queries past the end are zero (the softmax masks them).  One CTA per (tile index, head, batch).
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void split8_global(const float* v, __half* dst_hi, __half* dst_lo) {
    uint32_t h[4], l[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) split_h2(v[2 * e], v[2 * e + 1], h[e], l[e]);
    *reinterpret_cast<uint4*>(dst_hi) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(dst_lo) = make_uint4(l[0], l[1], l[2], l[3]);
}

template <int DQ, int DV>
__global__ void __launch_bounds__(256) attn_pack_kernel(const FAParams p, int nqt, int nblk) {
    using C = FTCfg<DQ, DV>;
    pdl_launch_dependents();
    pdl_wait();
    const int x = blockIdx.x, head = blockIdx.y, b = blockIdx.z;
    const size_t bh = (size_t)b * gridDim.y + head;
    const int S = p.T + p.Tx;
    unsigned char* wsb = reinterpret_cast<unsigned char*>(p.ws);
    const size_t q_all = (size_t)gridDim.z * gridDim.y * nqt * C::Q_BYTES;
    auto part = [&](const float* s1, int ld1, const float* s2, int ld2, size_t tok, int c) {
        return c < p.d1 ? s1 + tok * ld1 + head * p.d1 + c : s2 + tok * ld2 + head * p.d2 + (c - p.d1);
    };
    if (x < nqt) {
        __half* img = reinterpret_cast<__half*>(wsb + (bh * nqt + x) * C::Q_BYTES);
        const float qscale = p.scale * 1.4426950408889634f;      // exp2 domain
        for (int u = threadIdx.x; u < FT_BQ * (DQ / 8); u += blockDim.x) {
            const int r = u % FT_BQ, g = u / FT_BQ;
            const int row = x * FT_BQ + r;
            float v[8];
            load8(row < p.T ? part(p.q1, p.ldq1, p.q2, p.ldq2, (size_t)b * p.T + row, g * 8) : nullptr, v, qscale);
            __half* d = img + ((size_t)g * FT_BQ + r) * 8;
            split8_global(v, d, d + C::Q_PLANE / 2);
        }
    }
    if (x < nblk) {
        unsigned char* blk = wsb + q_all + (bh * nblk + x) * (size_t)(C::K_STAGE + C::V_STAGE);
        __half* kimg = reinterpret_cast<__half*>(blk);
        __half* vimg = reinterpret_cast<__half*>(blk + C::K_STAGE);
        const int k0 = x * FT_BK;
        for (int u = threadIdx.x; u < FT_BK * (DQ / 8); u += blockDim.x) {
            const int r = u % FT_BK, g = u / FT_BK;
            const int tk = k0 + r;
            const float* src = nullptr;
            if (tk < p.T) src = part(p.k1, p.ldk1, p.k2, p.ldk2, (size_t)b * p.T + tk, g * 8);
            else if (tk < S) src = part(p.xk1, p.ldx, p.xk2, p.ldx, (size_t)b * p.Tx + (tk - p.T), g * 8);
            float v[8];
            load8(src, v, 1.f);
            __half* d = kimg + ((size_t)g * FT_BK + r) * 8;
            split8_global(v, d, d + C::K_PLANE / 2);
        }
        // V^T: thread -> (group of 8 keys, quad of value dims): 8 row loads of 16 bytes, 4 hi + 4 lo units out
        for (int u = threadIdx.x; u < (FT_BK / 8) * (DV / 4); u += blockDim.x) {
            const int kg = u / (DV / 4), dq = u % (DV / 4);
            float4 vv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const int tk = k0 + kg * 8 + i;
                vv[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (tk < p.T) vv[i] = *reinterpret_cast<const float4*>(p.v + ((size_t)b * p.T + tk) * p.ldv + head * DV + dq * 4);
                else if (tk < S) vv[i] = *reinterpret_cast<const float4*>(p.xv + ((size_t)b * p.Tx + (tk - p.T)) * p.ldx + head * DV + dq * 4);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                float v[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) v[i] = e == 0 ? vv[i].x : (e == 1 ? vv[i].y : (e == 2 ? vv[i].z : vv[i].w));
                __half* d = vimg + ((size_t)kg * C::VROWS + dq * 4 + e) * 8;
                split8_global(v, d, d + DV * 8);
            }
        }
    }
}

template <int DQ, int DV>
__global__ void __launch_bounds__(FT_THREADS, 1) flash_attn_tc_kernel(const FAParams p) {
    using C = FTCfg<DQ, DV>;
    extern __shared__ __align__(128) unsigned char ft_smem[];
    const uint32_t sbase = smem_u32(ft_smem);
    const uint32_t bar0 = sbase + C::OFF_BAR;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ft_smem + C::OFF_BAR + 176);
    __nv_bfloat16* s_x = reinterpret_cast<__nv_bfloat16*>(ft_smem + C::OFF_X);   // [2 (block parity)][FT_NP (column part)][128 rows]
    const uint32_t Q_FULL = bar0, P_FULL = bar0 + 8;
#define KF(s) (bar0 + 16u + 8u * (s))
#define KE(s) (bar0 + 32u + 8u * (s))
#define VF(s) (bar0 + 48u + 8u * (s))
#define VE(s) (bar0 + 64u + 8u * (s))
#define SF(s) (bar0 + 80u + 8u * (s))
#define SE(s) (bar0 + 96u + 8u * (s))
#define OF(s) (bar0 + 112u + 8u * (s))
#define OE(s) (bar0 + 128u + 8u * (s))
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int q0 = blockIdx.x * FT_BQ, head = blockIdx.y, b = blockIdx.z;
    const int S = p.T + p.Tx;
    const int nblk = (S + FT_BK - 1) / FT_BK;

    if (threadIdx.x == 0) {
        mbar_init(Q_FULL, 1);
        mbar_init(P_FULL, FT_SOFTMAX);
        for (int s = 0; s < 2; ++s) {
            mbar_init(KF(s), 1); mbar_init(KE(s), 1);       // FULL: expect_tx of the producer + the copy's bytes
            mbar_init(VF(s), 1); mbar_init(VE(s), 1);
            mbar_init(SF(s), 1);          mbar_init(SE(s), FT_SOFTMAX);
            mbar_init(OF(s), 1);          mbar_init(OE(s), FT_SOFTMAX);
        }
        fence_barrier_init();
    }
    pdl_launch_dependents();
    if (warp == FT_WM) tmem_alloc(smem_u32(tmem_slot), C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    if (warp == FT_WP) {
        // ------------------------------------ producer: TMA-engine bulk copies of the packed tile images ------------------------------------
        // (attn_pack_kernel wrote Q / K / V^T of every (batch, head) as the exact shared-memory images, one contiguous block
        //  per tile: a stage is ONE cp.async.bulk, no conversion and no load latency chain inside this kernel)
        if (lane == 0) {
            unsigned long long* dbg = g_attn_dbg;
            unsigned long long dbg_acc[16] = {};
            const long long t_start = dbg ? clock64() : 0;
            const int nqt = gridDim.x;
            const size_t bh = (size_t)b * gridDim.y + head;
            const unsigned char* wsb = reinterpret_cast<const unsigned char*>(p.ws);
            const unsigned char* q_img = wsb + (bh * nqt + blockIdx.x) * C::Q_BYTES;
            const size_t q_all = (size_t)gridDim.z * gridDim.y * nqt * C::Q_BYTES;
            const unsigned char* k_img = wsb + q_all + bh * nblk * (size_t)(C::K_STAGE + C::V_STAGE);
            mbar_expect_tx(Q_FULL, C::Q_BYTES);
            bulk_copy_g2s(sbase + C::OFF_Q, q_img, C::Q_BYTES, Q_FULL);
            for (int j = 0; j < nblk; ++j) {
                const int s = j & 1;
                const uint32_t par = (uint32_t)(j >> 1) & 1u;
                const unsigned char* img = k_img + (size_t)j * (C::K_STAGE + C::V_STAGE);
                { FT_T0(); mbar_wait_quiet(KE(s), par ^ 1u); FT_ACC(11); }
                mbar_expect_tx(KF(s), C::K_STAGE);
                bulk_copy_g2s(sbase + C::OFF_K + s * C::K_STAGE, img, C::K_STAGE, KF(s));
                { FT_T0(); mbar_wait_quiet(VE(s), par ^ 1u); FT_ACC(12); }
                mbar_expect_tx(VF(s), C::V_STAGE);
                bulk_copy_g2s(sbase + C::OFF_V + s * C::V_STAGE, img + C::K_STAGE, C::V_STAGE, VF(s));
            }
            if (dbg) {
                unsigned long long* d = dbg + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16;
                d[10] = clock64() - t_start; d[11] = dbg_acc[11];
            }
        }
    } else if (warp == FT_WM) {
        // ------------------------------------ MMA issuer ------------------------------------
        constexpr uint32_t idesc_s = make_idesc_f16(128, FT_BK);
        constexpr uint32_t idesc_o2 = make_idesc_f16(128, 2 * DV), idesc_o1 = make_idesc_f16(128, DV);
        const uint32_t q_lo = desc_lo(sbase + C::OFF_Q, C::QSLAB);
        const uint32_t p_lo = desc_lo(sbase + C::OFF_P, C::PSLAB);
        unsigned long long* dbg = lane == 0 ? g_attn_dbg : nullptr;
        unsigned long long dbg_acc[16] = {};
        const long long t_start = dbg ? clock64() : 0;
        auto issue_qk = [&](int j) {
            const int s = j & 1;
            const uint32_t par = (uint32_t)(j >> 1) & 1u;
            { FT_T0(); mbar_wait_quiet(KF(s), par); FT_ACC(1); }
            { FT_T0(); mbar_wait_quiet(SE(s), par ^ 1u); FT_ACC(2); }
            tc_fence_after();
            if (elect_one()) {
                const uint32_t k_lo = desc_lo(sbase + C::OFF_K + s * C::K_STAGE, C::KSLAB);
                const uint32_t d = tmem_base + s * C::S_COLS;
#pragma unroll
                for (int ks = 0; ks < DQ / 16; ++ks) {
                    const uint32_t qa = q_lo + ((ks * 2 * C::QSLAB) >> 4), kb = k_lo + ((ks * 2 * C::KSLAB) >> 4);
                    tc_mma_f16_lh(d, qa, kb, idesc_s, ks != 0 ? 1u : 0u);                          // q_hi k_hi
                    tc_mma_f16_lh(d, qa + (C::Q_PLANE >> 4), kb, idesc_s, 1u);                     // q_lo k_hi
                    tc_mma_f16_lh(d, qa, kb + (C::K_PLANE >> 4), idesc_s, 1u);                     // q_hi k_lo
                }
                tc_commit(SF(s));
                tc_commit(KE(s));
            }
            __syncwarp();
        };
        mbar_wait_quiet(Q_FULL, 0);
        issue_qk(0);
        for (int j = 0; j < nblk; ++j) {
            if (j + 1 < nblk) issue_qk(j + 1);          // the tensor pipe works on S_{j+1} while the softmax warps turn S_j into P_j
            const int s = j & 1;
            const uint32_t par = (uint32_t)(j >> 1) & 1u;
            { FT_T0(); mbar_wait_quiet(P_FULL, (uint32_t)j & 1u); FT_ACC(3); }
            { FT_T0(); mbar_wait_quiet(VF(s), par); FT_ACC(4); }
            { FT_T0(); mbar_wait_quiet(OE(s), par ^ 1u); FT_ACC(5); }
            tc_fence_after();
            if (elect_one()) {
                const uint32_t v_lo = desc_lo(sbase + C::OFF_V + s * C::V_STAGE, C::VSLAB);
                const uint32_t d = tmem_base + 2 * C::S_COLS + s * C::O_COLS;
#pragma unroll
                for (int ks = 0; ks < FT_BK / 16; ++ks) {
                    const uint32_t pa = p_lo + ((ks * 2 * C::PSLAB) >> 4), vb = v_lo + ((ks * 2 * C::VSLAB) >> 4);
                    tc_mma_f16_lh(d, pa, vb, idesc_o2, ks != 0 ? 1u : 0u);                         // p_hi [v_hi ; v_lo]
                    tc_mma_f16_lh(d, pa + (C::P_PLANE >> 4), vb, idesc_o1, 1u);                    // p_lo v_hi
                }
                tc_commit(OF(s));
                tc_commit(VE(s));
            }
            __syncwarp();
        }
        if (dbg) {
            unsigned long long* d = dbg + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16;
            d[0] = clock64() - t_start;
            for (int i = 1; i <= 5; ++i) d[i] = dbg_acc[i];
        }
    } else {
        // ------------------------------------ softmax / output ------------------------------------
        // FT_NP threads per query row: warps w, w + 4, ... read the same TMEM lanes; part q owns key columns [q * HC, +HC) of a
        // block and value dims [q * HD, +HD).  (Four warps per scheduler: with two, the dependent exp2 / convert / store chains
        // left the softmax at ~3000 cycles per block against ~1500 of MMAs.)  Row maxima are exchanged through shared memory
        // once per block (named barrier 1 over the softmax threads); each thread keeps its own partial row sum.
        constexpr int HC = FT_BK / FT_NP, HD = DV / FT_NP;
        static_assert(HC == 32 && (HD == 8 || HD == 16), "softmax partition");
        const int part = warp >> 2;
        const int row = (warp & 3) * 32 + lane;        // 0..127 = TMEM lane
        const uint32_t lane_off = ((uint32_t)((warp & 3) * 32) << 16);
        float o[HD];
#pragma unroll
        for (int n = 0; n < HD; ++n) o[n] = 0.f;
        float m = -INFINITY, l = 0.f;
        unsigned long long* dbg = threadIdx.x == 0 ? g_attn_dbg : nullptr;
        unsigned long long dbg_acc[16] = {};
        const long long t_start = dbg ? clock64() : 0;
        auto fold_o = [&](int j) {                     // o += O_j: columns [part * HD, +HD) of the v_hi half and of the v_lo half
            const int s = j & 1;
            { FT_T0(); mbar_wait_quiet(OF(s), (uint32_t)(j >> 1) & 1u); FT_ACC(8); }
            tc_fence_after();
            const uint32_t t_o = tmem_base + lane_off + 2 * C::S_COLS + s * C::O_COLS + part * HD;
            float v[2 * HD];
            if (HD == 8) {
                tmem_ld_32x8_nowait(t_o, v);
                tmem_ld_32x8_nowait(t_o + DV, v + HD);
            } else {
                tmem_ld_32x16_nowait(t_o, v);
                tmem_ld_32x16_nowait(t_o + DV, v + HD);
            }
            tmem_ld_wait();
            tc_fence_before();
            mbar_arrive(OE(s));
#pragma unroll
            for (int i = 0; i < HD; ++i) o[i] += v[i] + v[HD + i];
        };
        for (int j = 0; j < nblk; ++j) {
            const int s = j & 1;
            const int k0 = j * FT_BK + part * HC;
            { FT_T0(); mbar_wait_quiet(SF(s), (uint32_t)(j >> 1) & 1u); FT_ACC(7); }
            tc_fence_after();
            float v[HC];
            const long long ts0 = dbg ? clock64() : 0;
            tmem_ld_32x32(tmem_base + lane_off + s * C::S_COLS + part * HC, v);
            tc_fence_before();
            mbar_arrive(SE(s));                        // the row lives in registers now: S_j may be overwritten by S_{j+2}
            if (dbg) dbg_acc[13] += clock64() - ts0;
            const bool tail = k0 + HC > S;
            float mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < HC; ++i) {
                if (tail && k0 + i >= S) v[i] = -INFINITY;
                mx = fmaxf(mx, v[i]);
            }
            // the reference point of the exponentials only has to be the SAME for every thread of the row and >= the true
            // maximum: the partial maxima travel as bf16 rounded UP (2 KB of shared memory instead of 4), and every thread
            // -- the owner included -- uses the rounded values
            __nv_bfloat16* xm = s_x + (j & 1) * FT_NP * FT_BQ;
            xm[part * FT_BQ + row] = __float2bfloat16_ru(mx);
            { FT_T0(); named_bar_sync(1, FT_SOFTMAX); FT_ACC(9); }
            float mn = m;
#pragma unroll
            for (int q = 0; q < FT_NP; ++q) mn = fmaxf(mn, __bfloat162float(xm[q * FT_BQ + row]));
            const float alpha = ex2_f(m - mn);         // exp2(-inf) = 0 on the first block
            m = mn;
            float rs = 0.f;
            const long long ts1 = dbg ? clock64() : 0;
#pragma unroll
            for (int i = 0; i < HC; ++i) {
                v[i] = ex2_f(v[i] - mn);               // masked columns: exp2(-inf) = 0
                rs += v[i];
            }
            l = l * alpha + rs;
            if (dbg) dbg_acc[14] += clock64() - ts1;
            // P shared memory / O accumulator of the previous block are free once its P V MMAs have retired
            if (j > 0) fold_o(j - 1);
            const long long ts2 = dbg ? clock64() : 0;
#pragma unroll
            for (int n = 0; n < HD; ++n) o[n] *= alpha;
#pragma unroll
            for (int g = 0; g < HC / 8; ++g) {
                const uint32_t a_hi = sbase + C::OFF_P + (part * (HC / 8) + g) * C::PSLAB + row * 16;
                split8_store(v + g * 8, a_hi, a_hi + C::P_PLANE);
            }
            const long long ts3 = dbg ? clock64() : 0;
            fence_proxy_async();
            mbar_arrive(P_FULL);
            if (dbg) { dbg_acc[15] += ts3 - ts2; dbg_acc[12] += clock64() - ts3; }
        }
        fold_o(nblk - 1);
        if (dbg) {
            unsigned long long* d = dbg + ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 16;
            d[6] = clock64() - t_start; d[7] = dbg_acc[7]; d[8] = dbg_acc[8]; d[9] = dbg_acc[9];
            d[13] = dbg_acc[13]; d[14] = dbg_acc[14]; d[15] = dbg_acc[15]; d[12] = dbg_acc[12];     // S load | exp | P store | fence + arrive
        }
        // ---- combine the partial row sums, normalise and store (conv operand layout, common.cuh) ----
        float* xl = reinterpret_cast<float*>(s_x);      // (every thread is past its last read of the maxima: fold_o waited for O)
        xl[part * FT_BQ + row] = l;
        named_bar_sync(1, FT_SOFTMAX);
        float lt = 0.f;
#pragma unroll
        for (int q = 0; q < FT_NP; ++q) lt += xl[q * FT_BQ + row];
        const float inv = 1.f / lt;
        const int tq = q0 + row;
        if (tq < p.T) {
            const int hh = tq / p.Wimg, ww = tq - hh * p.Wimg;
#pragma unroll
            for (int n = 0; n < HD; ++n)
                store_operand_elem(p.out, p.lo_off, p.parts, (size_t)b * (p.T / p.Wimg) + hh, p.C, p.Wimg, ww,
                                   head * DV + part * HD + n, o[n] * inv);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == FT_WM) {
        __syncwarp();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
#undef KF
#undef KE
#undef VF
#undef VE
#undef SF
#undef SE
#undef OF
#undef OE
}

template <int DQ, int DV>
static int launch_fa(const FAParams& p, int B, int heads, cudaStream_t st) {
    // B200_FA_IMPL = tc (default: tcgen05 / TMEM kernel) | mma (mma.sync register-level kernel: cross-check / A-B timing)
    static int impl = -1;
    if (impl < 0) {
        const char* e = getenv("B200_FA_IMPL");
        impl = (e && e[0] == 'm') ? 1 : 0;
    }
    if (impl == 0) {
        using C = FTCfg<DQ, DV>;
        static bool attr_tc = false;
        if (!attr_tc) {
            cudaError_t e = cudaFuncSetAttribute(flash_attn_tc_kernel<DQ, DV>, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
            if (e != cudaSuccess) {
                set_error("flash_attn_tc: cudaFuncSetAttribute(%d B) failed: %s", C::SMEM, cudaGetErrorString(e));
                return B200_E_CUDA;
            }
            attr_tc = true;
        }
        if (p.ws == nullptr) {
            set_error("flash_attention: the tcgen05 path needs a workspace of b200_flash_attention_workspace() bytes");
            return B200_E_ARG;
        }
        const int nqt = cdiv(p.T, FT_BQ), nblk = cdiv(p.T + p.Tx, FT_BK);
        launch_pdl(attn_pack_kernel<DQ, DV>, dim3(nqt > nblk ? nqt : nblk, heads, B), dim3(256), (size_t)0, st, p, nqt, nblk);
        B200_CHECK_LAUNCH();
        launch_pdl(flash_attn_tc_kernel<DQ, DV>, dim3(nqt, heads, B), dim3(FT_THREADS), (size_t)C::SMEM, st, p);
        B200_CHECK_LAUNCH();
        return B200_OK;
    }
    if (impl == 1) {
        const size_t smem = ((size_t)2 * FM_BK * (DQ + 8) + (size_t)2 * DV * (FM_BK + 8)) * sizeof(__half);
        dim3 grid(cdiv(p.T, FM_BQ), heads, B);
        launch_pdl(flash_attn_mma_kernel<DQ, DV>, grid, dim3(FM_THREADS), smem, st, p);
        B200_CHECK_LAUNCH();
        return B200_OK;
    }
    set_error("flash_attention: unknown B200_FA_IMPL");
    return B200_E_ARG;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_attn_set_debug(void* dbg_u64) {
    unsigned long long* p = (unsigned long long*)dbg_u64;
    cudaError_t e = cudaMemcpyToSymbol(b200::g_attn_dbg, &p, sizeof(p));
    if (e != cudaSuccess) {
        b200::set_error("attn_set_debug: %s", cudaGetErrorString(e));
        return B200_E_CUDA;
    }
    return B200_OK;
}

extern "C" size_t b200_flash_attention_workspace(int B, int heads, int T, int Tx, int dq, int dv) {
    const size_t nqt = (size_t)cdiv(T, FT_BQ), nblk = (size_t)cdiv(T + Tx, FT_BK);
    return (size_t)B * heads * (nqt * FT_BQ * dq * 4 + nblk * FT_BK * (size_t)(dq + dv) * 4);
}

extern "C" int b200_flash_attention(const float* qkv, int E, void* out, int out_w, int parts, int B, int heads, int T,
                                    float scale, void* workspace, void* stream) {
    // self-attention on the fused in-projection output qkv fp32 [B,T,3E] (q | k | v, head-major)
    B200_CHECK_ARG(qkv && out && (parts >= 1 && parts <= 3) && heads > 0 && E % heads == 0);
    B200_CHECK_ARG(out_w > 0 && out_w % OTW == 0 && T % out_w == 0);
    const int d = E / heads;
    FAParams p{qkv, nullptr, qkv + E, nullptr, qkv + 2 * E, nullptr, nullptr, nullptr, 3 * E, 0, 3 * E, 0, 3 * E, 0,
               d, 0, d, T, 0, (__half*)out, (size_t)B * T / out_w * (out_w / OTW) * (E / 8) * OPX * 8, parts, E, out_w, scale,
               (__half*)workspace};
    if (d == 64) return launch_fa<64, 64>(p, B, heads, (cudaStream_t)stream);
    if (d == 32) return launch_fa<32, 32>(p, B, heads, (cudaStream_t)stream);
    set_error("flash_attention: head dim %d not supported (32 or 64)", d);
    return B200_E_ARG;
}

extern "C" int b200_flash_attention_oa(const float* qkv, const float* pos_p, const float* kl, const float* pos_l,
                                       const float* vl, void* out, int out_w, int parts, int B, int C, int heads, int T,
                                       int L2, float scale2, void* workspace, void* stream) {
    B200_CHECK_ARG(qkv && pos_p && kl && pos_l && vl && out && (parts >= 1 && parts <= 3));
    B200_CHECK_ARG(heads > 0 && C % heads == 0 && C / heads == 32 && out_w > 0 && out_w % OTW == 0 && T % out_w == 0 && L2 >= 0);
    FAParams p{qkv, pos_p, qkv + C, pos_p, qkv + 2 * C, kl, pos_l, vl, 3 * C, C, 3 * C, C, 3 * C, C,
               32, 32, 32, T, L2, (__half*)out, (size_t)B * T / out_w * (out_w / OTW) * (C / 8) * OPX * 8, parts, C, out_w, scale2,
               (__half*)workspace};
    return launch_fa<64, 32>(p, B, heads, (cudaStream_t)stream);
}
