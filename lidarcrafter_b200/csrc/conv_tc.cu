// K1: ring-padded 3x3 / 1x1 convolution as an implicit GEMM on the sm_100a tensor cores.
//
//   out[b,h,w,n] = ( w_inv * sum_{dy,dx,k} a[b, h+dy-1, (w+dx-1) mod W, k] * wgt[n,k,dy,dx] + bias[n] + res[b,h,w,n] ) * scale
//
// replaces F.pad(circular)+nn.Conv2d (+bias, +skip, *1/sqrt2, next GroupNorm statistics) of the reference
// (lidargen/models/unets/ops.py:32-49,149-173; efficient_unet.py:104-115).
//
// Design: PERSISTENT CTAs (one per SM) loop over output tiles; one tile = 128 consecutive pixels of R image rows x
// BN output channels.
//   * operands fp16, accumulation fp32 in TMEM.  Two accumulator sets of R x (128 lanes x BN columns) so the
//     epilogue of tile i overlaps the MMAs of tile i+1.
//   * precision: NP = 1 -> one fp16 MMA per product (rel. error ~2e-3 through the whole UNet);
//                NP = 2 -> error-compensated split a = a_hi + a_lo, w = w_hi + w_lo (all fp16):
//                a*w ~= a_hi*w_hi + a_lo*w_hi + a_hi*w_lo (3 MMAs, ~2^-22 relative) -- the mode that meets the
//                reference's 1e-3 fp32 tolerance.  Weights are pre-scaled by a power of two (w_inv undoes it).
//                NP = 3 -> fp16 main term + fp8 correction: the two small cross terms a_lo*w_hi + a_hi*w_lo (2^-11 of
//                the result) only need ~3 significant bits, so they are evaluated by ONE kind::f8f6f4 MMA (K = 32 =
//                [e4m3(a_lo 2^11) | e4m3(a)] x [e4m3(w 2^-11) | e4m3(w_lo)]) at twice the fp16 rate: 2 MMA time
//                units per product instead of 3, ~5e-5 relative through the whole UNet.  (NP = 4: same operands,
//                corrections in their own accumulator columns, added in the epilogue.)
//   * A (activations) live in HBM "tile-major": [plane][b][h][W/128][C/8][130][8] fp16 (common.cuh): for one image
//     row, one 128-pixel tile and one 8-channel group, 130 pixels (tile + its two ring-halo pixels) at a 16-byte
//     pitch -- which IS the canonical no-swizzle K-major shared-memory layout of a tcgen05 operand, halo included.
//     For every K chunk the R+2 halo rows are staged ONCE by the TMA engine, one cp.async.bulk per (row, plane)
//     (zero rows from a zero page), and each of the 9 filter taps is just a different START ADDRESS of the same
//     slab (dx*16 bytes, dy = other row): no im2col copies, no 9x re-reads, no swizzle.  (Measured on B200: the
//     MMA rate does not depend on the 16-byte alignment of the start address -- profiles/r01_mma_probe.txt -- but
//     the TMA engine is request-rate bound, hence few large copies.)
//   * B (weights): host-prepacked tiles in exactly the shared-memory image, one cp.async.bulk per (chunk, filter row).
//   * warp roles: 0-3 epilogue (TMEM -> regs -> smem transpose -> coalesced fp32 store + residual + per-channel
//     sum / sum-of-squares for the next GroupNorm), 4 activation producer and 6 weight producer (TMA bulk copies,
//     mbarrier tx bytes; independent warps so a full weight ring never stalls the activation prefetch),
//     5 MMA issuer (single thread, tcgen05.mma kind::f16, M=128 N=BN K=16) + TMEM owner.
#include <stdlib.h>

#include <type_traits>

#include "common.cuh"

namespace b200 {

struct ConvParams {
    const __half* a;
    const __half* w;
    const float* bias;
    const float* res;
    float* out;
    double* stats;
    float out_scale;
    float w_inv;  // 1 / (power-of-two scale applied to the packed weights)
    int B, H, W, Cin, Cout, ring;
    int n_tiles;
    // fused front end (b200_conv_gn_tc, template FUSE): the A operand is produced IN the kernel from the fp32 NHWC
    // activation(s) x0 [B,H*W,C0] (| x1 [B,H*W,C1]: channel concat) -- GroupNorm(+AdaGN)-apply with the complete per-channel
    // statistics st0 / st1 (fp64 {sum, sumsq} [B,C,2]; nullptr: no normalisation), optional SiLU, fp16 hi/lo (or e4m3 pair)
    // split -- by four transform warps that write the K-major shared-memory slab the MMA issuers read.
    const float* x0;
    const float* x1;
    const double* st0;
    const double* st1;
    const float* gn_gamma;
    const float* gn_beta;
    const float* gn_ada;
    int C0, C1, gn_ada_stride, gn_groups, gn_silu;
    float gn_eps;
    // split-K (b200_conv_tc_splitk): blockIdx.y = K slice; slice s accumulates the chunks [s, s + 1) * (Cin / KC / k_splits) and
    // writes its raw partial sums to out + s * split_stride (no bias / residual / statistics: conv_splitk_reduce_kernel)
    int k_splits;
    long long split_stride;
};

__device__ __align__(128) unsigned char g_zero_page[16384];  // source of zero-padding rows / pixels

// Optional in-kernel profile (b200_conv_set_debug): per CTA 8 x u64 cycle counters
//   [0] MMA thread total, [1] wait FULL_A, [2] wait FULL_B, [3] wait ACC_EMPTY,
//   [4] epilogue warp 0 total, [5] epilogue wait ACC_FULL, [6] producer wait EMPTY_A, [7] producer wait EMPTY_B
__device__ unsigned long long* g_conv_dbg = nullptr;
// Timing ablations (b200_conv_set_ablate; results are WRONG with any bit set -- bottleneck analysis only):
//   1 MMA thread does not wait for FULL_A, 2 ... for FULL_B, 4 epilogue skips the residual loads / global stores,
//   8 no tcgen05.mma is issued (commits only), 16 producers arrive without copying (no TMA traffic),
//   32 the issuing warp idles ~300 cycles after every chunk (is its own time hidden behind queued MMAs?)
__device__ int g_conv_ablate = 0;
#ifdef B200_CONV_ABLATE          // diagnostic builds only (make EXTRA=-DB200_CONV_ABLATE); the product build compiles them out
#define ABL(bit) (ablate & (bit))
#else
#define ABL(bit) false
#endif
#define DBG_T0() const long long t0__ = dbg ? clock64() : 0
#define DBG_ACC(slot) do { if (dbg) dbg_acc[slot] += clock64() - t0__; } while (0)

constexpr int PIX = 128;  // pixels per tile row (= MMA M)
constexpr int CONV_THREADS = 384;   // warps 0-3 and 8-11: epilogue (two per TMEM lane quarter); 4: A producer; 5, 7: MMA; 6: B producer
constexpr int FUSE_THREADS = 640;   // fused front end: + warps 12-19 = transform warps (they replace the A producer)
constexpr int XF_WARPS = 8;
constexpr int MAX_CIN = 1024;       // per-channel GroupNorm coefficients kept in shared memory by the fused front end
constexpr int EPI_WARPS_MAX = 8;   // warps 0-3 and 8-11; ConvCfg::EW of them work (BN = 128 tiles: 4, see ConvCfg)

template <int BN, int R, int TAPS, int NP, bool FUSE = false>
struct ConvCfg {
    static_assert(!FUSE || NP == 2 || NP == 3, "fused front end: fp16x3 / fp16f8 operands only");
    static constexpr int THREADS = FUSE ? FUSE_THREADS : CONV_THREADS;
    static constexpr int COEF = FUSE ? 2 * 4 * MAX_CIN + 256 : 0;   // s_a[MAX_CIN], s_b[MAX_CIN], {mean, rstd}[32]
    static constexpr int PL = NP == 1 ? 1 : 2;    // operand planes (hi | lo or fp8 pair)
    static constexpr bool F8 = NP >= 3;
    static constexpr bool SEP = NP == 4;
    static constexpr int KC = NP == 1 ? 32 : 16;  // channels per K chunk
    static constexpr int KG = KC / 8;             // 8-channel groups (slabs) per row
    static constexpr int KS = KC / 16;            // MMA K steps per chunk
    static constexpr int HALO = TAPS == 9 ? 1 : 0;   // halo ROWS above / below the tile
    static constexpr int DX0 = TAPS == 9 ? 0 : 1;    // first pixel position read by tap dx = 0 (1x1: the body starts at 1)
    static constexpr int NPX = OPX;                  // 128 pixels + the two halo pixels of the operand layout
    static constexpr int SLAB = NPX * 16;  // bytes: one 8-channel group of one staged row
    static constexpr int RA = R + 2 * HALO;
    static constexpr int NSLAB = PL * RA * KG;
    static constexpr int A_PART = RA * KG * SLAB;
    static constexpr int A_STAGE = PL * A_PART;
    static constexpr int TW = TAPS == 9 ? 3 : 1;  // taps per B stage (one filter row)
    static constexpr int TG = TAPS == 9 ? 3 : 1;  // B stages per chunk
    static constexpr int B_PART = BN * KC * 2;
    static constexpr int B_TAP = PL * B_PART;
    static constexpr int B_STAGE = TW * B_TAP;
    // Epilogue warps: 8 (two per TMEM lane quarter) for BN = 64, where the full-resolution layers keep the epilogue ~50 % busy;
    // 4 for BN = 128, whose epilogue is 75-99 % idle: the 22 KB of staging it gives back buys two more weight stages, and
    // the weight ring must hold TWO chunks (the waiting issuer polls one chunk ahead of the one being multiplied).
    static constexpr int EW = BN == 128 ? 4 : 8;
    static constexpr int EG = EW / 4;                           // warps per TMEM lane quarter = item stride
    static constexpr int EPI_STG = EW * 32 * 36 * 4;            // per-warp transpose staging
    static constexpr int EPI = EPI_STG + EW * 2 * BN * 4;       // + per-warp channel sum / sum-of-squares partials
    static constexpr int BAR_BYTES = 448;                       // 4 rings x 12 mbarriers + 4 accumulator barriers + TMEM slot
    static constexpr int BUDGET = 227 * 1024 - EPI - BAR_BYTES - COEF;
#ifndef B200_CONV_SA_MAX
#define B200_CONV_SA_MAX 3
#endif
    static constexpr int TGC = TAPS == 9 ? 3 : 1;               // weight stages per chunk
    // Issue unit of an MMA issuer warp = GRP consecutive K chunks.  A 1x1 conv has only R (x2-3) MMAs per 16-channel chunk
    // (~200-600 cycles of tensor pipe) against ~1000 cycles of per-visit work of the issuer, so its chunks are visited in
    // groups; their stages are small, the rings just get deeper (2 groups in flight + 1 chunk of slack).
    static constexpr int GRP = TAPS == 1 ? (R == 1 ? 4 : (R == 2 ? 2 : 1)) : 1;
    // 3x3: 3 activation chunks in flight if five weight stages still fit next to them, else 2
    static constexpr int SA = TAPS == 1 ? 2 * GRP + 1
                                        : ((BUDGET - (2 * TGC - 1) * B_STAGE) / A_STAGE >= B200_CONV_SA_MAX ? B200_CONV_SA_MAX : 2);
    static_assert(SA <= 12, "barrier block holds 12 slots per ring");
    static constexpr int SB_RAW = (BUDGET - SA * A_STAGE) / B_STAGE;
    static constexpr int SB = SB_RAW > 12 ? 12 : SB_RAW;   // weight ring: as deep as shared memory allows
    static_assert(SB >= 3, "weight ring too shallow");
    static constexpr int OFF_A = 0;
    static constexpr int OFF_B = SA * A_STAGE;
    static constexpr int OFF_EPI = OFF_B + SB * B_STAGE;
    static constexpr int OFF_COEF = OFF_EPI + EPI;
    static constexpr int OFF_BAR = OFF_COEF + COEF;
    static constexpr int SMEM = OFF_BAR + BAR_BYTES;
    // MERGE (fp16x3 whenever 2 x R x BN accumulator columns fit twice in TMEM): the weight tile keeps hi and lo rows
    // adjacent ([KG][hi|lo][BN][8]) so that a_hi x [w_hi ; w_lo] is ONE N = 2 BN MMA (A is fetched from shared memory
    // once for 2 BN accumulator columns instead of twice) and only a_lo x w_hi remains an N = BN MMA.  The two partial
    // accumulators (columns [0,BN) and [BN,2BN)) are added in the epilogue.
    static constexpr bool MERGE = (NP == 2 && 2 * R * BN <= 256);
    static constexpr int ACC_ROW = (MERGE || SEP) ? 2 * BN : BN;
    static constexpr int ACC_COLS = R * ACC_ROW;
    static constexpr int TMEM_COLS = 2 * ACC_COLS;
    static_assert(TMEM_COLS >= 32 && TMEM_COLS <= 512 && (TMEM_COLS & (TMEM_COLS - 1)) == 0, "TMEM columns");
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
    static constexpr int ROWB = KG * SLAB;   // bytes of one staged (row, plane): ONE bulk copy
    static_assert(ROWB <= 16384, "zero page too small");
};

// ---------------------------------------------------------------------------------------------------------
// fused front end (FUSE): GroupNorm(+AdaGN)-apply + SiLU + fp16 hi/lo (or e4m3 pair) split, straight into the A ring
// ---------------------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void reg_alloc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N>
__device__ __forceinline__ void reg_dealloc() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// read-only activation load that does not allocate in the (small: 256 KB - 227 KB of shared memory) L1
__device__ __forceinline__ float4 ldg_stream_f4(const float* ptr) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];"
                 : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
                 : "l"(ptr));
    return v;
}
__device__ __forceinline__ void sts_v2(uint32_t addr, uint32_t a, uint32_t b) {
    asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(addr), "r"(a), "r"(b) : "memory");
}
__device__ __forceinline__ void sts_b32(uint32_t addr, uint32_t a) {
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(addr), "r"(a) : "memory");
}

// per-channel affine coefficients of GroupNorm(+AdaGN) for sample b: y = x * s_a[c] + s_b[c] (same fp32 expression order
// as gn_act_kernel).  Run by the COEFFICIENT WARP (warp 4 of the fused variant, the idle TMA-producer slot) -- not by the
// transform warps: their unit loop keeps a stage of prefetched activations in registers and must not contain calls or
// fp64 code.  One lane per group (<= 32 groups): the 32 group statistics are reduced in parallel (the first version
// walked the groups serially with a warp reduction each: ~3 us in front of the first operand stage of every launch).
// Handshake: COEF_FULL (1 arrival: coefficients of the next sample are in shared memory) / COEF_EMPTY (XF_WARPS arrivals:
// every transform warp is done with the current ones).
__device__ __forceinline__ void coef_warp(const ConvParams& p, float* s_a, float* s_b, float* s_mr, uint32_t bar_full,
                                          uint32_t bar_empty, int b_lo, int b_hi, int lane) {
    const int Ctot = p.Cin;
    for (int b = b_lo, k = 0; b <= b_hi; ++b, ++k) {
        mbar_wait(bar_empty, (k & 1) ^ 1);
        if (p.st0 != nullptr) {
            const int cpg = Ctot / p.gn_groups;
            if (lane < p.gn_groups) {
                const int gi = lane;
                double su = 0.0, ss = 0.0;
                for (int i = 0; i < cpg; ++i) {
                    const int c = gi * cpg + i;
                    const double2 st = *reinterpret_cast<const double2*>(
                        c < p.C0 ? p.st0 + ((size_t)b * p.C0 + c) * 2 : p.st1 + ((size_t)b * p.C1 + (c - p.C0)) * 2);
                    su += st.x;
                    ss += st.y;
                }
                const double n = (double)p.H * p.W * cpg;
                const double mean = su / n;
                double var = ss / n - mean * mean;
                if (var < 0.0) var = 0.0;
                s_mr[2 * gi] = (float)mean;
                s_mr[2 * gi + 1] = (float)(1.0 / sqrt(var + (double)p.gn_eps));
            }
            __syncwarp();
            for (int c = lane; c < Ctot; c += 32) {
                const int gi = c / cpg;
                float a = s_mr[2 * gi + 1], bb = -s_mr[2 * gi] * s_mr[2 * gi + 1];
                float ga = 1.f, be = 0.f, sc = 1.f, sh = 0.f;
                if (p.gn_gamma) { ga = p.gn_gamma[c]; be = p.gn_beta[c]; }
                if (p.gn_ada) {
                    sc = 1.f + p.gn_ada[(size_t)b * p.gn_ada_stride + c];
                    sh = p.gn_ada[(size_t)b * p.gn_ada_stride + Ctot + c];
                }
                a *= ga; bb = bb * ga + be;
                a *= sc; bb = bb * sc + sh;
                s_a[c] = a;
                s_b[c] = bb;
            }
        } else {
            for (int c = lane; c < Ctot; c += 32) { s_a[c] = 1.f; s_b[c] = 0.f; }
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(bar_full);
    }
}

// Eight transform warps (two per SM sub-partition) produce every A stage = one 16-channel K chunk of the RA = R + 2 halo
// rows of a tile.  The elementwise work (GroupNorm-apply, SiLU, split: ~10 instructions per element, two of them MUFU) on
// halo-duplicated rows is 30-40 % of the SM's issue slots while the tensor pipe runs a stage, so the role is written for
// instruction count AND instruction-level parallelism (two transform warps per scheduler cannot hide a ~200-cycle
// FFMA -> EX2 -> RCP -> F2FP -> STS chain per unit by multithreading: the first version, one unit at a time with a
// branch per unit, needed ~3.5 us per stage against ~1 us of MMAs):
//   * unit = 8 pixels x 16 channels = ONE 16-byte load per lane: lane -> (8-channel group g = lane / 16, pixel p8 =
//     (lane / 2) % 8, channel quad q = lane % 2); a warp-level load covers 8 pixels x 64 contiguous bytes; per pixel one
//     8-byte store into the fp16 hi slab and one 8-byte (lo) or two 4-byte (L8, A8) stores into plane 1 -- every
//     warp-level store covers 128 contiguous bytes of a slab (conflict-free);
//   * warp w owns the pixel groups w and w + 8 of EVERY staged row: 2 RA units per stage with compile-time (row, group)
//     -> all shared-memory offsets are immediates on one per-lane base;
//   * units are processed in branch-free groups of XF_ILP (straight-line code: the compiler interleaves the chains);
//     rows outside the image (first / last tile row only) take a second copy of the loop that selects zeros;
//   * (3x3) the 8 RA ring-halo quads of a stage (RA rows x 2 sides x 4 channel quads) are ONE more unit; the warps take
//     it in turns (stage i: warp i % 8), so a warp does 8 1/8 units per stage on average instead of 9;
//   * memory-level parallelism: the units of stage s + 1 are loaded into the register slots of stage s as these are
//     consumed (one whole stage, 32 KB per SM, in flight), across tile boundaries too.
constexpr int XF_ILP = 4;

template <int NP, bool SILU, bool SEL>
__device__ __forceinline__ void xf_store(const float4 x4, const float (&ca)[4], const float (&cb)[4], bool valid,
                                         uint32_t a_hi, uint32_t a_p1, uint32_t a_p1b, int ablate = 0) {
    float y[4] = {fmaf(x4.x, ca[0], cb[0]), fmaf(x4.y, ca[1], cb[1]), fmaf(x4.z, ca[2], cb[2]), fmaf(x4.w, ca[3], cb[3])};
    if (SILU) {
#pragma unroll
        for (int e = 0; e < 4; ++e) y[e] = silu_f(y[e]);
    }
    const __half2 h01 = __floats2half2_rn(y[0], y[1]), h23 = __floats2half2_rn(y[2], y[3]);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const float lo[4] = {y[0] - f01.x, y[1] - f01.y, y[2] - f23.x, y[3] - f23.y};
    uint32_t w0 = *reinterpret_cast<const uint32_t*>(&h01), w1 = *reinterpret_cast<const uint32_t*>(&h23), w2, w3;
    if (NP == 2) {
        const __half2 l01 = __floats2half2_rn(lo[0], lo[1]), l23 = __floats2half2_rn(lo[2], lo[3]);
        w2 = *reinterpret_cast<const uint32_t*>(&l01);
        w3 = *reinterpret_cast<const uint32_t*>(&l23);
    } else {
        // plane 1 of a 16-channel chunk: slab 0 = L8 = e4m3(lo * 2^11), slab 1 = A8 = e4m3(x), 16 bytes per pixel each
        w2 = f8x4(lo[0] * F8_LO_SCALE, lo[1] * F8_LO_SCALE, lo[2] * F8_LO_SCALE, lo[3] * F8_LO_SCALE);
        w3 = f8x4(y[0], y[1], y[2], y[3]);
    }
    if (SEL) {   // zero padding (rows outside the image, non-ring edges) is exact: zeros, not act(b)
        w0 = valid ? w0 : 0u; w1 = valid ? w1 : 0u; w2 = valid ? w2 : 0u; w3 = valid ? w3 : 0u;
    }
    if (ABL(256)) return;
    if (ABL(1024) && (w0 ^ w1 ^ w2 ^ w3) != 0x5eadbeefu) return;    // keeps the math alive, drops the stores
    sts_v2(a_hi, w0, w1);
    if (NP == 2) {
        sts_v2(a_p1, w2, w3);
    } else {
        sts_b32(a_p1, w2);
        sts_b32(a_p1b, w3);
    }
}

template <int BN, int R, int TAPS, int NP, bool SILU>
__device__ __forceinline__ void xform_warps(const ConvParams& p, uint8_t* smem, uint32_t sbase, uint32_t bar0, int tw,
                                            int lane, int tile_lo, int tile_hi, int ablate) {
    using C = ConvCfg<BN, R, TAPS, NP, true>;
    constexpr int RA = C::RA, HALO = C::HALO, KC = C::KC;
    constexpr int UPS = 2 * RA;                 // body units per warp and stage
    (void)ablate;
    unsigned long long* dbg = (tw == 0 && lane == 0) ? g_conv_dbg : nullptr;
    unsigned long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_start = dbg ? clock64() : 0;
    constexpr bool HAS_HALO = TAPS == 9;
    constexpr int HQ = 8 * RA;                  // ring-halo quads of a stage: RA rows x 2 sides x 4 channel quads
    constexpr int HWN = (HQ + 31) / 32;         // warps on halo duty per stage (1; RA = 6: 2)
    static_assert(KC == 16 && PIX == 16 * XF_WARPS && HWN <= XF_WARPS, "fused front end geometry");
    const float* s_a = reinterpret_cast<const float*>(smem + C::OFF_COEF);
    const float* s_b = s_a + MAX_CIN;
    const uint32_t bar_coef_full = bar0 + 424u, bar_coef_empty = bar0 + 432u;
    const int g = lane >> 4, p8 = (lane >> 1) & 7, q = lane & 1;
    const int co = g * 8 + q * 4;                           // this lane's 4 channels inside the 16-channel chunk
    const int WT = p.W / PIX, HG = p.H / R, NT = p.Cout / BN, NCH = p.Cin / KC;
    const int px0 = tw * 8 + p8;                            // this lane's pixel inside the tile (first group; second: + 64)
    // shared-memory byte offsets of this lane inside a stage: unit (r, e) adds r * ROWB + e * 64 * 16
    const uint32_t so_hi = sbase + C::OFF_A + g * C::SLAB + (1 + px0) * 16 + q * 8;
    const uint32_t so_p1 = sbase + C::OFF_A + C::A_PART + (NP == 2 ? g * C::SLAB + (1 + px0) * 16 + q * 8 : (1 + px0) * 16 + co);

    // position in the stage stream: tile coordinates (order: wt fastest, then hg, n-tile, sample) + K chunk
    struct Cur { int tile, c, wt, hg, nt, b; };
    auto advance = [&](Cur& cu) {
        if (++cu.c == NCH) {
            cu.c = 0;
            ++cu.tile;
            if (++cu.wt == WT) {
                cu.wt = 0;
                if (++cu.hg == HG) {
                    cu.hg = 0;
                    if (++cu.nt == NT) { cu.nt = 0; ++cu.b; }
                }
            }
        }
    };
    // bit r: staged row r of the tile row group hg lies inside the image
    auto row_mask = [&](int hg) {
        uint32_t m = 0;
        const int h0 = hg * R - HALO;
#pragma unroll
        for (int r = 0; r < RA; ++r) m |= ((unsigned)(h0 + r) < (unsigned)p.H ? 1u : 0u) << r;
        return m;
    };
    // fp32 source of K chunk c: tensor x0 | x1 (channel concat), its channel count, first channel inside it
    auto chunk_src = [&](int c, const float*& src, int& Cs) {
        const int cbase = c * KC;
        if (cbase < p.C0) { src = p.x0 + cbase; Cs = p.C0; }
        else { src = p.x1 + (cbase - p.C0); Cs = p.C1; }
    };
    float4 ring[UPS + 1];         // [UPS]: the halo unit, when this warp is on halo duty for the stage
#pragma unroll
    for (int u = 0; u <= UPS; ++u) ring[u] = make_float4(0.f, 0.f, 0.f, 0.f);

    // issue the loads of a whole stage `cu` (warp-uniform position) into the ring; ia_of = its index in this CTA's stream
    uint32_t l_mask = 0;
    const char* l_ptr = nullptr;         // this lane's element of staged row 0, pixel group 0
    uint32_t l_row = 0, l_e = 0;         // byte strides: image row, 64 pixels
    const float* l_hptr = nullptr;       // this lane's ring-halo quad (halo duty only)
    bool l_h = false;
    auto load_ctx = [&](const Cur& cu, uint32_t ia_of) {
        const bool on = cu.tile < tile_hi;
        l_mask = on ? row_mask(cu.hg) : 0u;
        const float* src;
        int Cs;
        chunk_src(cu.c, src, Cs);
        l_row = (uint32_t)(p.W * Cs * 4);
        l_e = (uint32_t)(64 * Cs * 4);
        const int h0 = cu.hg * R - HALO;
        l_ptr = reinterpret_cast<const char*>(src) +
                (((long long)(cu.b * p.H + h0) * p.W + cu.wt * PIX + px0) * Cs + co) * 4;
        if (HAS_HALO) {
            const uint32_t k = ((uint32_t)tw - ia_of * HWN) & (XF_WARPS - 1);      // halo slot of this warp in that stage
            const int idx = (int)k * 32 + lane;
            const int row = idx >> 3, side = (idx >> 2) & 1, cq = idx & 3;
            int ww = cu.wt * PIX + (side ? PIX : -1);
            bool ok = on && k < HWN && idx < HQ && ((l_mask >> row) & 1u);
            if (ww < 0) { ww += p.W; ok = ok && p.ring; }
            else if (ww >= p.W) { ww -= p.W; ok = ok && p.ring; }
            l_h = ok;
            l_hptr = src + ((long long)(cu.b * p.H + h0 + row) * p.W + ww) * Cs + cq * 4;
        }
    };
    auto load_unit = [&](int u) {          // u: compile-time
        if (u < UPS) {
            const int r = u >> 1, e = u & 1;
            if (((l_mask >> r) & 1u) && !ABL(64))        // warp-uniform
                ring[u] = ldg_stream_f4(reinterpret_cast<const float*>(l_ptr + (size_t)(r * l_row + e * l_e)));
        } else if (l_h) {
            ring[UPS] = ldg_stream_f4(l_hptr);
        }
    };

    if (tile_lo >= tile_hi) return;
    Cur cp{tile_lo, 0, 0, 0, 0, 0};
    {
        int t = tile_lo;
        cp.wt = t % WT; t /= WT;
        cp.hg = t % HG; t /= HG;
        cp.nt = t % NT;
        cp.b = t / NT;
    }
    uint32_t ia = 0, n_coef = 0;
    load_ctx(cp, 0);
#pragma unroll
    for (int u = 0; u <= UPS; ++u)
        if (u < UPS || HAS_HALO) load_unit(u);

    int cur_b = -1;
    while (cp.tile < tile_hi) {
        const uint32_t s = ia % C::SA;
        if (cp.b != cur_b) {         // (tiles of a CTA are contiguous: a few sample changes per launch at most)
            if (cur_b >= 0) {        // this warp holds no more coefficients of the previous sample
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_coef_empty);
            }
            mbar_wait_quiet(bar_coef_full, n_coef & 1);
            ++n_coef;
            cur_b = cp.b;
        }
        float ca[4], cb[4];
        {
            const float4 a4 = *reinterpret_cast<const float4*>(s_a + cp.c * KC + co);
            const float4 b4 = *reinterpret_cast<const float4*>(s_b + cp.c * KC + co);
            ca[0] = a4.x; ca[1] = a4.y; ca[2] = a4.z; ca[3] = a4.w;
            cb[0] = b4.x; cb[1] = b4.y; cb[2] = b4.z; cb[3] = b4.w;
        }
        const uint32_t vm = row_mask(cp.hg);
        const uint32_t st_hi = so_hi + s * C::A_STAGE, st_p1 = so_p1 + s * C::A_STAGE;
        Cur nx = cp;
        advance(nx);
        {
            DBG_T0();
            mbar_wait_quiet(bar0 + 96u + 8u * s, ((ia / C::SA) & 1) ^ 1);       // EMPTY_A(s)
            DBG_ACC(6);
        }
        load_ctx(nx, ia + 1);                        // the NEXT stage: its units take over the register slots as they free up
        auto body = [&](auto sel_tag) {
            constexpr bool SEL = decltype(sel_tag)::value;
#pragma unroll
            for (int u0 = 0; u0 < UPS; u0 += XF_ILP) {
                float4 x[XF_ILP];
#pragma unroll
                for (int j = 0; j < XF_ILP; ++j)
                    if (u0 + j < UPS) {
                        x[j] = ring[u0 + j];
                        load_unit(u0 + j);
                    }
#pragma unroll
                for (int j = 0; j < XF_ILP; ++j)
                    if (u0 + j < UPS) {
                        const int r = (u0 + j) >> 1, e = (u0 + j) & 1;
                        const uint32_t off = r * C::ROWB + e * 64 * 16;
                        xf_store<NP, SILU, SEL>(x[j], ca, cb, (vm >> r) & 1u, st_hi + off, st_p1 + off, st_p1 + off + C::SLAB, ablate);
                    }
            }
        };
#ifdef B200_XF_PROF
        const long long tb0 = clock64();
#endif
        if (ABL(128)) {}
        else if (vm == (1u << RA) - 1u) body(std::false_type{});
        else body(std::true_type{});
#ifdef B200_XF_PROF
        const long long tb1 = clock64();
        dbg_acc[1] += tb1 - tb0;
#endif
        if (HAS_HALO) {
            const uint32_t k = ((uint32_t)tw - ia * HWN) & (XF_WARPS - 1);
            const float4 x4 = ring[UPS];
            load_unit(UPS);                          // (only if this warp is on halo duty for the next stage)
            if (k < HWN) {                           // warp-uniform: this warp converts the stage's ring-halo quads
                const int idx = (int)k * 32 + lane;
                const int row = idx >> 3, side = (idx >> 2) & 1, cq = idx & 3;
                const int ww = cp.wt * PIX + (side ? PIX : -1);
                const bool valid = ((vm >> row) & 1u) && (p.ring || (ww >= 0 && ww < p.W));
                float ha[4], hb[4];
                {
                    const float4 a4 = *reinterpret_cast<const float4*>(s_a + cp.c * KC + cq * 4);
                    const float4 b4 = *reinterpret_cast<const float4*>(s_b + cp.c * KC + cq * 4);
                    ha[0] = a4.x; ha[1] = a4.y; ha[2] = a4.z; ha[3] = a4.w;
                    hb[0] = b4.x; hb[1] = b4.y; hb[2] = b4.z; hb[3] = b4.w;
                }
                const uint32_t pos = (side ? OPX - 1 : 0) * 16;
                const uint32_t stage = sbase + C::OFF_A + s * C::A_STAGE + row * C::ROWB;
                const uint32_t h_hi = stage + (cq >> 1) * C::SLAB + pos + (cq & 1) * 8;
                const uint32_t h_p1 = stage + C::A_PART + (NP == 2 ? (cq >> 1) * C::SLAB + pos + (cq & 1) * 8 : pos + cq * 4);
                if (idx < HQ) xf_store<NP, SILU, true>(x4, ha, hb, valid, h_hi, h_p1, h_p1 + C::SLAB);
            }
        }
#ifdef B200_XF_PROF
        const long long tb2 = clock64();
        dbg_acc[2] += tb2 - tb1;
#endif
        if (!ABL(512)) fence_proxy_async();          // generic-proxy stores -> visible to the tensor core's async-proxy operand reads
        __syncwarp();
        if (lane == 0) mbar_arrive(bar0 + 8u * s);                    // FULL_A(s): one arrival per transform warp
#ifdef B200_XF_PROF
        dbg_acc[3] += clock64() - tb2;
#endif
        ++ia;
        cp = nx;
    }
    if (dbg) {
        dbg[blockIdx.x * 8 + 6] = dbg_acc[6];
        dbg[blockIdx.x * 8 + 7] = clock64() - t_start;
#ifdef B200_XF_PROF
        dbg[blockIdx.x * 8 + 1] = dbg_acc[1];
        dbg[blockIdx.x * 8 + 2] = dbg_acc[2];
        dbg[blockIdx.x * 8 + 3] = dbg_acc[3];
#endif
    }
}

template <int BN, int R, int TAPS, int NP, bool FUSE>
__global__ void __launch_bounds__(FUSE ? FUSE_THREADS : CONV_THREADS, 1) conv_tc_kernel(const ConvParams p) {
    using C = ConvCfg<BN, R, TAPS, NP, FUSE>;
    constexpr int KC = C::KC;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + C::OFF_BAR;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::OFF_BAR + 416);
#define FULL_A(s) (bar0 + 8u * (s))
#define EMPTY_A(s) (bar0 + 96u + 8u * (s))
#define FULL_B(s) (bar0 + 192u + 8u * (s))
#define EMPTY_B(s) (bar0 + 288u + 8u * (s))
#define ACC_FULL(s) (bar0 + 384u + 8u * (s))
#define ACC_EMPTY(s) (bar0 + 400u + 8u * (s))

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int NT = p.Cout / BN, WT = p.W / PIX, HG = p.H / R;
    const int NCH_ALL = p.Cin / KC;
    const int NCH = NCH_ALL / (p.k_splits > 1 ? p.k_splits : 1);     // K chunks of this CTA (split-K: its slice)
    const int cb = (int)blockIdx.y * NCH;                            // first chunk of the slice
    const int CG = p.Cin / 8;
    // contiguous chunk of tiles per CTA (same image rows / same batch index: L2 locality, few statistic flushes)
    const int tiles_per_cta = (p.n_tiles + gridDim.x - 1) / gridDim.x;
    const int tile_lo = blockIdx.x * tiles_per_cta;
    const int tile_hi = min(tile_lo + tiles_per_cta, p.n_tiles);

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::SA; ++s) {
            mbar_init(FULL_A(s), FUSE ? XF_WARPS : 1);   // fused: one arrival per transform warp; else the expect_tx of the TMA producer
            mbar_init(EMPTY_A(s), 1);
        }
        for (int s = 0; s < C::SB; ++s) {
            mbar_init(FULL_B(s), 1);
            mbar_init(EMPTY_B(s), 1);
        }
        for (int s = 0; s < 2; ++s) {
            mbar_init(ACC_FULL(s), 2);     // one commit from each of the two MMA issuer warps
            mbar_init(ACC_EMPTY(s), C::EW * 32);
        }
        if (FUSE) {
            mbar_init(bar0 + 424u, 1);             // COEF_FULL
            mbar_init(bar0 + 432u, XF_WARPS);      // COEF_EMPTY
        }
        fence_barrier_init();
    }
    pdl_launch_dependents();   // let the next kernel's CTAs be scheduled while this grid drains
    if (warp == 5) tmem_alloc(smem_u32(tmem_slot), C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // everything above (barriers, TMEM) overlapped the previous kernel's tail.  (Letting the weight producer skip this wait
    // -- packed weights are constants of a plan -- measured no gain and would be wrong when a conv directly follows the
    // kernel that packs its weights, as in the tile autotune and the kernel tests.)
    pdl_wait();

    const int ablate = g_conv_ablate;
    (void)ablate;
    // Role dispatch by WARPGROUP first: the fused variant (640 threads: 96 registers per thread at launch) re-balances the
    // register file with setmaxnreg -- epilogue warpgroups 120 (32 accumulator + 32 residual prefetch registers; the
    // next-item residual prefetch of the 384-thread variant is dropped), issuers / weight producer / coefficient warp 56,
    // transform warpgroups 88 (one stage of prefetched units + coefficients) = 472 of the 480 x 128 registers.  Every warp
    // of a warpgroup must execute the same setmaxnreg, and it must dominate the code that uses the registers.
    const int wg = warp >> 2;
    if (wg >= 3) {
        if constexpr (FUSE) {
            reg_dealloc<88>();
            if (p.gn_silu) xform_warps<BN, R, TAPS, NP, true>(p, smem, sbase, bar0, warp - 12, lane, tile_lo, tile_hi, ablate);
            else xform_warps<BN, R, TAPS, NP, false>(p, smem, sbase, bar0, warp - 12, lane, tile_lo, tile_hi, ablate);
        }
    } else if (wg == 1) {
    if constexpr (FUSE) reg_dealloc<56>();
    if (warp == 4 && FUSE) {
        // ------------------------------ coefficient warp (fused front end) ------------------------------
        if (tile_lo < tile_hi) {
            float* s_a = reinterpret_cast<float*>(smem + C::OFF_COEF);
            const int tps = WT * HG * NT;        // tiles per sample
            coef_warp(p, s_a, s_a + MAX_CIN, s_a + 2 * MAX_CIN, bar0 + 424u, bar0 + 432u, tile_lo / tps,
                      (tile_hi - 1) / tps, lane);
        }
    } else if (warp == 4) {
        // ------------------------------ producer warp: TMA-engine bulk copies for A and B ------------------------------
        uint32_t ia = 0;
        unsigned long long* dbg = g_conv_dbg;
        unsigned long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const size_t part_elems = (size_t)p.B * p.H * WT * CG * OPX * 8;
        for (int tile = tile_lo; tile < tile_hi; ++tile) {
            int t = tile;
            const int wt = t % WT; t /= WT;
            const int hg = t % HG; t /= HG;
            const int b = t / NT;
            const int h0 = hg * R;
            for (int c = 0; c < NCH; ++c) {
                {
                    const int s = ia % C::SA;
                    const uint32_t ph = (ia / C::SA) & 1;
                    if (lane == 0) {
                        DBG_T0();
                        mbar_wait(EMPTY_A(s), ph ^ 1);
                        DBG_ACC(6);
                        if (ABL(16)) mbar_arrive(FULL_A(s));
                        else mbar_expect_tx(FULL_A(s), C::A_STAGE);
                    }
                    __syncwarp();
                    if (ABL(16)) { ++ia; continue; }
                    const uint32_t dst0 = sbase + C::OFF_A + s * C::A_STAGE;
                    // one bulk copy per (plane, staged row): the KG slabs of this K chunk are contiguous in the
                    // tile-major operand and already carry the two halo pixels
                    for (int r2 = lane; r2 < C::PL * C::RA; r2 += 32) {
                        const int part = r2 / C::RA;
                        const int r = r2 - part * C::RA;
                        const int gh = h0 + r - C::HALO;
                        const uint32_t dst = dst0 + r2 * C::ROWB;
                        if (gh < 0 || gh >= p.H) {
                            bulk_copy_g2s(dst, g_zero_page, C::ROWB, FULL_A(s));
                            continue;
                        }
                        const __half* src = p.a + part * part_elems +
                                            operand_unit((size_t)b * p.H + gh, WT, CG, wt, (cb + c) * C::KG, 0) * 8;
                        const bool edge_l = !p.ring && wt == 0, edge_r = !p.ring && wt == WT - 1;
                        if (!(edge_l || edge_r)) {
                            bulk_copy_g2s(dst, src, C::ROWB, FULL_A(s));
                        } else {
                            // zero padding in W (not on the hot path): per slab, zero halo pixel(s) + the rest
                            for (int j = 0; j < C::KG; ++j) {
                                const uint32_t d = dst + j * C::SLAB;
                                const __half* sj = src + (size_t)j * OPX * 8;
                                const int lo_px = edge_l ? 1 : 0, hi_px = edge_r ? OPX - 1 : OPX;
                                if (edge_l) bulk_copy_g2s(d, g_zero_page, 16, FULL_A(s));
                                bulk_copy_g2s(d + lo_px * 16, sj + lo_px * 8, (hi_px - lo_px) * 16, FULL_A(s));
                                if (edge_r) bulk_copy_g2s(d + (OPX - 1) * 16, g_zero_page, 16, FULL_A(s));
                            }
                        }
                    }
                    ++ia;
                }
                __syncwarp();
            }
        }
        if (dbg && lane == 0) dbg[blockIdx.x * 8 + 6] = dbg_acc[6];
    } else if (warp == 6) {
        // ------------------------------ weight producer: one thread, TMA-engine bulk copies ------------------------------
        if (lane == 0) {
            uint32_t ib = 0;
            unsigned long long* dbg = g_conv_dbg;
            unsigned long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int tile = tile_lo; tile < tile_hi; ++tile) {
                const int nt = (tile / (WT * HG)) % NT;
                const __half* wsrc = p.w + ((size_t)nt * NCH_ALL + cb) * TAPS * (C::PL * BN * KC);
                for (int q = 0; q < NCH * C::TG; ++q, ++ib) {
                    const int s = ib % C::SB;
                    const uint32_t ph = (ib / C::SB) & 1;
                    {
                        DBG_T0();
                        mbar_wait(EMPTY_B(s), ph ^ 1);
                        DBG_ACC(7);
                    }
                    if (ABL(16)) { mbar_arrive(FULL_B(s)); continue; }
                    mbar_expect_tx(FULL_B(s), C::B_STAGE);
                    bulk_copy_g2s(sbase + C::OFF_B + s * C::B_STAGE, wsrc + (size_t)q * (C::B_STAGE / 2), C::B_STAGE,
                                  FULL_B(s));
                }
            }
            if (dbg) dbg[blockIdx.x * 8 + 7] = dbg_acc[7];
        }
    } else if (warp == 5 || warp == 7) {
        // ------------------------------ MMA issuers (two warps, ping-pong over the K chunks) ------------------------------
        // Measured with the ablation build (tools/exp_conv_ablate.py): the tensor pipe queues only a couple of MMAs, so every
        // cycle the issuing warp spends NOT issuing (barrier polls ~140 cycles each, fence, election, descriptor moves into
        // uniform registers, commits: 800-1300 cycles per chunk against ~2000 cycles of MMAs) is a cycle the pipe idles --
        // an artificial 300-cycle pause per chunk lengthened the kernel by exactly 300 cycles x chunks.  Hence two issuers:
        // while one warp's MMAs of issue unit g (a K chunk; a group of chunks for 1x1 convs) run, the other has already
        // waited for the operands of unit g + 1 and built
        // its descriptors, and starts issuing the moment it is handed the turn (named barriers 2 / 3, ~tens of cycles).
        // Each warp runs its loop with uniform control flow (waits included) and one elected lane issues: the descriptor
        // arithmetic stays in uniform registers (one 32-bit add per operand between two tcgen05.mma).  Every warp commits
        // the stages it read and, per tile, its own share of the accumulator (ACC_FULL counts two arrivals).
        {
            const uint32_t par = warp == 7 ? 1u : 0u;
            constexpr uint32_t idesc = make_idesc_f16(128, BN);
            constexpr uint32_t idesc2 = make_idesc_f16(128, 2 * BN);
            constexpr uint32_t KGS = C::MERGE ? 2 * BN * 16 : BN * 16;   // byte stride between 8-channel groups of B
            uint32_t ia = 0, ib = 0, it = 0, ig = 0;     // chunk / weight-stage / tile / issue-unit counters
            unsigned long long* dbg = (lane == 0 && par == 0) ? g_conv_dbg : nullptr;
            unsigned long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const long long t_start = dbg ? clock64() : 0;
            for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
                const uint32_t buf = it & 1;
                {
                    DBG_T0();
                    mbar_wait(ACC_EMPTY(buf), ((it >> 1) & 1) ^ 1);
                    DBG_ACC(3);
                }
                tc_fence_after();
                const uint32_t acc = tmem_base + buf * C::ACC_COLS;
                for (int c = 0; c < NCH; c += C::GRP, ++ig) {
                    const int ng = min(C::GRP, NCH - c);     // chunks of this issue unit (warp-uniform)
                    if ((ig & 1u) != par) {                  // the other issuer's unit
                        ia += ng;
                        ib += ng * C::TG;
                        continue;
                    }
                    // (all 32 lanes poll: a lane-0-only poll + __syncwarp() was measured 25 % SLOWER -- the divergent region
                    // takes the descriptor arithmetic out of the uniform datapath)
                    if (!ABL(1)) {
                        DBG_T0();
#pragma unroll
                        for (int g = 0; g < C::GRP; ++g)
                            if (g < ng) mbar_wait(FULL_A((ia + g) % C::SA), ((ia + g) / C::SA) & 1);
                        DBG_ACC(1);
                    }
                    if (!ABL(2)) {
                        DBG_T0();
#pragma unroll
                        for (int q = 0; q < C::GRP * C::TG; ++q)
                            if (q < ng * C::TG) mbar_wait(FULL_B((ib + q) % C::SB), ((ib + q) / C::SB) & 1);
                        DBG_ACC(2);
                    }
                    uint32_t a_lo[C::GRP], b_lo[C::GRP * C::TG];
#pragma unroll
                    for (int g = 0; g < C::GRP; ++g) a_lo[g] = desc_lo(sbase + C::OFF_A + ((ia + g) % C::SA) * C::A_STAGE, C::SLAB);
#pragma unroll
                    for (int q = 0; q < C::GRP * C::TG; ++q) b_lo[q] = desc_lo(sbase + C::OFF_B + ((ib + q) % C::SB) * C::B_STAGE, KGS);
                    if (ABL(32)) {   // is this warp's own time hidden now?  (+300 cycles of preparation per unit)
                        const long long t_spin = clock64();
                        while (clock64() - t_spin < 300) {}
                    }
                    if (ig != 0) {   // my turn: the other issuer has issued unit ig - 1
                        DBG_T0();
                        named_bar_sync(2 + par, 64);
                        DBG_ACC(4);
                    }
                    tc_fence_after();
                    if (elect_one()) {
#pragma unroll
                        for (int g = 0; g < C::GRP; ++g) {
                            if (g < ng) {
#pragma unroll
                                for (int dy = 0; dy < C::TG; ++dy) {
                                    const uint32_t first_cd = (uint32_t)(((c + g) | dy) != 0);
                                    if (!ABL(8))
#pragma unroll
                                    for (int dx = 0; dx < C::TW; ++dx) {
#pragma unroll
                                        for (int o = 0; o < R; ++o) {
#pragma unroll
                                            for (int ks = 0; ks < C::KS; ++ks) {
                                                // staged input row o + dy feeds output row o through filter row dy
                                                const uint32_t a_hi = a_lo[g] + ((uint32_t)(dy * C::KG * C::SLAB) >> 4) +
                                                                      (((o * C::KG + ks * 2) * C::SLAB + (dx + C::DX0) * 16) >> 4);
                                                const uint32_t b_hi = b_lo[g * C::TG + dy] + ((dx * C::B_TAP + ks * 2 * KGS) >> 4);
                                                const uint32_t first = (dx | ks) != 0 ? 1u : first_cd;
                                                const uint32_t d = acc + o * C::ACC_ROW;
                                                if (C::F8) {
                                                    // plane 1 of both operands: [L8 | A8] x [e4m3(w 2^-11) | e4m3(w_lo)], K = 32
                                                    tc_mma_f16_lh(d, a_hi, b_hi, idesc, first);
                                                    tc_mma_f8_lh(d + (C::SEP ? BN : 0), a_hi + (C::A_PART >> 4),
                                                                 b_hi + (C::B_PART >> 4), idesc, C::SEP ? first : 1u);
                                                } else if (C::MERGE) {
                                                    tc_mma_f16_lh(d, a_hi, b_hi, idesc2, first);                  // [hi*hi | hi*lo]
                                                    tc_mma_f16_lh(d, a_hi + (C::A_PART >> 4), b_hi, idesc, 1u);   // += lo*hi
                                                } else {
                                                    tc_mma_f16_lh(d, a_hi, b_hi, idesc, first);
                                                    if (NP == 2) {
                                                        tc_mma_f16_lh(d, a_hi + (C::A_PART >> 4), b_hi, idesc, 1u);
                                                        tc_mma_f16_lh(d, a_hi, b_hi + (C::B_PART >> 4), idesc, 1u);
                                                    }
                                                }
                                            }
                                        }
                                    }
                                    tc_commit(EMPTY_B((ib + g * C::TG + dy) % C::SB));  // weights slot free once these MMAs retire
                                }
                                tc_commit(EMPTY_A((ia + g) % C::SA));
                            }
                        }
                    }
                    __syncwarp();
                    tc_fence_before();
                    named_bar_arrive(2 + (par ^ 1u), 64);    // hand the tensor pipe to the other issuer
                    ia += ng;
                    ib += ng * C::TG;
                }
                // this warp's share of the tile's accumulator is complete when ITS MMAs have retired
                if (elect_one()) tc_commit(ACC_FULL(buf));
                __syncwarp();
            }
            if (dbg) {
                dbg[blockIdx.x * 8 + 0] = clock64() - t_start;
#ifndef B200_XF_PROF
                dbg[blockIdx.x * 8 + 1] = dbg_acc[1];
                dbg[blockIdx.x * 8 + 2] = dbg_acc[2];
                dbg[blockIdx.x * 8 + 3] = dbg_acc[3];
#endif
            }
        }
    }
    } else {
    if constexpr (FUSE) reg_alloc<120>();
    if (warp < 4 || C::EW == 8) {
        // ------------------------------ epilogue: warps 0-3 and 8-11; warp w reads TMEM lanes 32*(w%4) .. +31 ------------------------------
        // the two warps of a lane quarter split the (row, 32-column slice) work items of a tile between them
        const int ew = warp < 4 ? warp : warp - 4;       // 0..7: epilogue warp index
        const int quarter = warp & 3, egroup = ew >> 2;  // TMEM lane quarter, work-item parity
        float* stg = reinterpret_cast<float*>(smem + C::OFF_EPI) + ew * (32 * 36);
        const int col4 = lane & 7, rb = lane >> 3;
        const float scale = p.out_scale, winv = p.w_inv;
        float* const out_base = p.out + (size_t)blockIdx.y * (size_t)p.split_stride;
        uint32_t it = 0;
        unsigned long long* dbg = g_conv_dbg;
        unsigned long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        const long long t_start = dbg ? clock64() : 0;
        float* sst = reinterpret_cast<float*>(smem + C::OFF_EPI + C::EPI_STG) + ew * (2 * BN);   // [2][BN] of this warp
        for (int i = lane; i < 2 * BN; i += 32) sst[i] = 0.f;
        __syncwarp();
        int cur_b = -1, cur_nt = -1;
        // flush the CTA's per-channel partial sums: ONE fp64 atomic pair per channel per (batch, n-tile) change instead of
        // one per warp x row x tile (same-address atomics serialise at L2: 1024 of them per address cost ~100 us)
        auto flush_stats = [&]() {
            named_bar_sync(1, C::EW * 32);
            if (cur_b >= 0) {
                float* all = reinterpret_cast<float*>(smem + C::OFF_EPI + C::EPI_STG);
                for (int ch = ew * 32 + lane; ch < BN; ch += C::EW * 32) {
                    float a = 0.f, q = 0.f;
#pragma unroll
                    for (int w = 0; w < C::EW; ++w) {
                        a += all[w * 2 * BN + ch];
                        q += all[w * 2 * BN + BN + ch];
                        all[w * 2 * BN + ch] = 0.f;
                        all[w * 2 * BN + BN + ch] = 0.f;
                    }
                    double* st = p.stats + ((size_t)cur_b * p.Cout + cur_nt * BN + ch) * 2;
                    atomicAdd(st, (double)a);
                    atomicAdd(st + 1, (double)q);
                }
            }
            named_bar_sync(1, C::EW * 32);
        };
        for (int tile = tile_lo; tile < tile_hi; ++tile, ++it) {
            int t = tile;
            const int wt = t % WT; t /= WT;
            const int hg = t % HG; t /= HG;
            const int nt = t % NT;
            const int b = t / NT;
            const int w0 = wt * PIX, h0 = hg * R, n0 = nt * BN;
            if (p.stats && (b != cur_b || nt != cur_nt)) {
                flush_stats();
                cur_b = b;
                cur_nt = nt;
            }
            const uint32_t buf = it & 1;
            constexpr int NSL = BN / 32, NITEM = R * NSL;
            // residual rows of an item: 8 x 16 bytes per lane, requested BEFORE the accumulator wait / the TMEM read of the
            // item so that the (HBM) latency of the skip tensor overlaps the MMAs instead of extending the epilogue
            auto res_ptr = [&](int item) {
                const int o = item / NSL, sl = item - o * NSL;
                return p.res + ((size_t)(b * p.H + h0 + o) * p.W + w0 + quarter * 32 + rb) * p.Cout + n0 + sl * 32 + col4 * 4;
            };
            float4 rv[8];
            const bool do_res = p.res && !ABL(4);
            if (do_res && egroup < NITEM) {
                const float* rp = res_ptr(egroup);
#pragma unroll
                for (int i = 0; i < 8; ++i) rv[i] = *reinterpret_cast<const float4*>(rp + (size_t)(4 * i) * p.Cout);
            }
            {
                DBG_T0();
                mbar_wait(ACC_FULL(buf), (it >> 1) & 1);
                DBG_ACC(5);
            }
            tc_fence_after();
            const uint32_t acc = tmem_base + buf * C::ACC_COLS + ((uint32_t)(quarter * 32) << 16);
            for (int item = egroup; item < NITEM; item += C::EG) {
                const int o = item / NSL, sl = item - o * NSL;
                const int h = h0 + o;
                const size_t row_base = ((size_t)(b * p.H + h) * p.W + w0 + quarter * 32) * p.Cout;
                {
                    float v[32];
                    tmem_ld_32x32(acc + o * C::ACC_ROW + sl * 32, v);
                    if (C::MERGE || C::SEP) {
                        float v2[32];
                        tmem_ld_32x32(acc + o * C::ACC_ROW + BN + sl * 32, v2);
#pragma unroll
                        for (int j = 0; j < 32; ++j) v[j] += v2[j];
                    }
                    float4 rn[8];      // next item's residual: in flight during this item's transpose / stores
                    if (!FUSE && do_res && item + C::EG < NITEM) {
                        const float* rp = res_ptr(item + C::EG);
#pragma unroll
                        for (int i = 0; i < 8; ++i) rn[i] = *reinterpret_cast<const float4*>(rp + (size_t)(4 * i) * p.Cout);
                    }
                    if (item + C::EG >= NITEM) {
                        // this warp's TMEM reads of the accumulator set are done -> hand it back to the MMA warp
                        tc_fence_before();
                        mbar_arrive(ACC_EMPTY(buf));
                    }
#pragma unroll
                    for (int j = 0; j < 8; ++j)
                        *reinterpret_cast<float4*>(stg + lane * 36 + 4 * j) =
                            make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
                    __syncwarp();
                    const int nb = n0 + sl * 32 + col4 * 4;
                    const float4 bi = p.bias ? *reinterpret_cast<const float4*>(p.bias + nb) : make_float4(0, 0, 0, 0);
                    float s1[4] = {0, 0, 0, 0}, s2[4] = {0, 0, 0, 0};
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const int row = rb + 4 * i;
                        float4 tv = *reinterpret_cast<const float4*>(stg + row * 36 + col4 * 4);
                        const size_t gi = row_base + (size_t)row * p.Cout + nb;
                        tv.x = fmaf(tv.x, winv, bi.x); tv.y = fmaf(tv.y, winv, bi.y);
                        tv.z = fmaf(tv.z, winv, bi.z); tv.w = fmaf(tv.w, winv, bi.w);
                        if (do_res) { tv.x += rv[i].x; tv.y += rv[i].y; tv.z += rv[i].z; tv.w += rv[i].w; }
                        tv.x *= scale; tv.y *= scale; tv.z *= scale; tv.w *= scale;
                        if (!ABL(4)) *reinterpret_cast<float4*>(out_base + gi) = tv;
                        s1[0] += tv.x; s1[1] += tv.y; s1[2] += tv.z; s1[3] += tv.w;
                        s2[0] += tv.x * tv.x; s2[1] += tv.y * tv.y; s2[2] += tv.z * tv.z; s2[3] += tv.w * tv.w;
                    }
                    if (FUSE && do_res && item + C::EG < NITEM) {
                        // 640-thread variant (120 registers here): the next item's residual is requested as soon as this
                        // item's has been consumed -- in flight during the statistics and the next TMEM read / transpose
                        const float* rp = res_ptr(item + C::EG);
#pragma unroll
                        for (int i = 0; i < 8; ++i) rv[i] = *reinterpret_cast<const float4*>(rp + (size_t)(4 * i) * p.Cout);
                    }
                    if (p.stats) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], 8);
                            s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], 16);
                            s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], 8);
                            s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], 16);
                        }
                        if (rb == 0) {   // 8 lanes x 4 channels = the 32 channels of this slice; warp-private rows
                            const int ch = sl * 32 + col4 * 4;
#pragma unroll
                            for (int e = 0; e < 4; ++e) {
                                sst[ch + e] += s1[e];
                                sst[BN + ch + e] += s2[e];
                            }
                        }
                    }
                    __syncwarp();
                    if (!FUSE && do_res && item + C::EG < NITEM) {
#pragma unroll
                        for (int i = 0; i < 8; ++i) rv[i] = rn[i];
                    }
                }
            }
        }
        if (p.stats) flush_stats();
        if (dbg && threadIdx.x == 0) {
            dbg[blockIdx.x * 8 + 4] = clock64() - t_start;
            dbg[blockIdx.x * 8 + 5] = dbg_acc[5];
        }
    }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        __syncwarp();
        tmem_dealloc(tmem_base, C::TMEM_COLS);
    }
#undef FULL_A
#undef EMPTY_A
#undef FULL_B
#undef EMPTY_B
#undef ACC_FULL
#undef ACC_EMPTY
}

template <int BN, int R, int TAPS, int NP, bool FUSE>
static int launch_conv(ConvParams p, int num_sms, cudaStream_t st) {
    using C = ConvCfg<BN, R, TAPS, NP, FUSE>;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, R, TAPS, NP, FUSE>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) {
            set_error("conv_tc: cudaFuncSetAttribute(%d B smem) failed: %s", C::SMEM, cudaGetErrorString(e));
            return B200_E_CUDA;
        }
        attr_set = true;
    }
    p.n_tiles = (p.W / PIX) * (p.H / R) * p.B * (p.Cout / BN);
    int grid = p.n_tiles < num_sms ? p.n_tiles : num_sms;
    const int tpc = (p.n_tiles + grid - 1) / grid;
    grid = (p.n_tiles + tpc - 1) / tpc;     // no empty CTAs with contiguous chunks
    launch_pdl_if(pdl_enabled_conv(), conv_tc_kernel<BN, R, TAPS, NP, FUSE>, dim3(grid, p.k_splits > 1 ? p.k_splits : 1),
                  dim3(C::THREADS), (size_t)C::SMEM, st, p);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

}  // namespace b200
#include "conv_col.cuh"
namespace b200 {

// ---------------------------------------------------------------------------------------------------------
// weight packing: OIHW fp32 -> fp16 tiles in the exact shared-memory image of a B stage
//   layout [Cout/bn][Cin/KC][taps][parts][KC/8][bn][8]; parts = 1 (fp16) or 2 (hi, lo); values pre-multiplied
//   by `wscale` (a power of two chosen on the host so max|w|*wscale is in [256, 512)).
// ---------------------------------------------------------------------------------------------------------
__global__ void pack_weight_kernel(const float* __restrict__ w, __half* __restrict__ out, int Cout, int Cin, int taps,
                                   int bn, int parts, int merged, float wscale) {
    const int kc = parts == 2 ? 16 : 32, kg = kc / 8;
    const size_t total = (size_t)Cout * Cin * taps * parts;
    const int nch = Cin / kc;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i;
        const int e = r % 8; r /= 8;
        const int n = r % bn; r /= bn;
        int j, part;
        if (merged) {   // merged layout [KG][hi|lo][bn][8] (see ConvCfg::MERGE)
            part = r % parts; r /= parts;
            j = r % kg; r /= kg;
        } else {
            j = r % kg; r /= kg;
            part = r % parts; r /= parts;
        }
        const int tap = r % taps; r /= taps;
        const int c = r % nch; r /= nch;
        const int nt = (int)r;
        const int k = c * kc + j * 8 + e;
        const int co = nt * bn + n;
        const float v = w[((size_t)co * Cin + k) * taps + tap] * wscale;
        const __half hi = __float2half_rn(v);
        out[i] = part == 0 ? hi : __float2half_rn(v - __half2float(hi));
    }
}

// parts = 3 ("fp16 + fp8 correction"): per (n-tile, 16-channel chunk, tap) the B stage image is
//   plane 0: [2 (8-channel groups)][bn][8] fp16(w*s)         plane 1: [2][bn][16] e4m3: {w*s*2^-11, w*s - fp16(w*s)}
// with s a power of two such that max|w*s| is in [2^14, 2^15).
__global__ void pack_weight_f8_kernel(const float* __restrict__ w, __half* __restrict__ out, int Cout, int Cin, int taps,
                                      int bn, float wscale) {
    const size_t total = (size_t)Cout * Cin * taps;
    const int nch = Cin / 16;
    uint8_t* out8 = reinterpret_cast<uint8_t*>(out);
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i;
        const int e16 = r % 16; r /= 16;
        const int n = r % bn; r /= bn;
        const int tap = r % taps; r /= taps;
        const int c = r % nch; r /= nch;
        const int nt = (int)r;
        const int k = c * 16 + e16, co = nt * bn + n;
        const float v = w[((size_t)co * Cin + k) * taps + tap] * wscale;
        const __half hi = __float2half_rn(v);
        const size_t tap_base = (((size_t)nt * nch + c) * taps + tap) * (size_t)(bn * 32);   // halves per tap image
        out[tap_base + ((size_t)(e16 / 8) * bn + n) * 8 + (e16 & 7)] = hi;
        const size_t p1 = (tap_base + (size_t)bn * 16) * 2;                                   // byte offset of plane 1
        out8[p1 + (size_t)n * 16 + e16] = f8x1(v * (1.f / F8_LO_SCALE));
        out8[p1 + ((size_t)bn + n) * 16 + e16] = f8x1(v - __half2float(hi));
    }
}

// plain layout for the cross-check kernel: [parts][tap][Cout][Cin]
__global__ void pack_weight_plain_kernel(const float* __restrict__ w, __half* __restrict__ out, int Cout, int Cin,
                                         int taps, int parts, float wscale) {
    const size_t per = (size_t)Cout * Cin * taps;
    for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < per * parts;
         i += (size_t)gridDim.x * blockDim.x) {
        size_t r = i % per;
        const int part = (int)(i / per);
        const int k = r % Cin; r /= Cin;
        const int co = r % Cout; r /= Cout;
        const int tap = (int)r;
        const float v = w[((size_t)co * Cin + k) * taps + tap] * wscale;
        const __half hi = __float2half_rn(v);
        if (parts == 3)   // fp16 images of the e4m3 values the tensor-core path multiplies with
            out[i] = part == 0 ? hi
                               : __float2half_rn(f8_to_float(f8x1(part == 1 ? v * (1.f / F8_LO_SCALE) : v - __half2float(hi))));
        else
            out[i] = part == 0 ? hi : __float2half_rn(v - __half2float(hi));
    }
}

// ---------------------------------------------------------------------------------------------------------
// CUDA-core cross-check implementation (same operands, same epilogue contract)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void load8h(const __half* p, float* v) {
    const uint4 raw = *reinterpret_cast<const uint4*>(p);
    const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h2[e]);
        v[2 * e] = f.x; v[2 * e + 1] = f.y;
    }
}

__global__ void __launch_bounds__(256) conv_ffma_kernel(const ConvParams p, int taps, int parts) {
    // block: 32 pixels (lanes) x 32 output channels (8 warps x 4)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long npix = (long long)p.B * p.H * p.W;
    const long long pix = (long long)blockIdx.x * 32 + lane;
    const int co0 = blockIdx.y * 32 + warp * 4;
    const bool valid = pix < npix;
    const int w = valid ? (int)(pix % p.W) : 0;
    const int h = valid ? (int)((pix / p.W) % p.H) : 0;
    const int b = valid ? (int)(pix / ((long long)p.W * p.H)) : 0;
    const int kside = taps == 9 ? 3 : 1, pad = taps == 9 ? 1 : 0;
    const size_t a_part = (size_t)p.B * p.H * (p.W / OTW) * (p.Cin / 8) * OPX * 8, w_part = (size_t)taps * p.Cout * p.Cin;
    float acc[4] = {0, 0, 0, 0};
    for (int tap = 0; tap < taps; ++tap) {
        const int dy = tap / kside - pad, dx = tap % kside - pad;
        const int gh = h + dy;
        int gw = w + dx;
        bool ok = valid && gh >= 0 && gh < p.H;
        if (gw < 0) { if (p.ring) gw += p.W; else ok = false; }
        else if (gw >= p.W) { if (p.ring) gw -= p.W; else ok = false; }
        // tile-major operand (common.cuh): [plane][b][h][W/128][C/8][130][8]; the body pixel gw sits at position gw%128 + 1
        const int WT = p.W / OTW, CG = p.Cin / 8;
        const size_t bh = (size_t)b * p.H + gh;
        const int twt = gw / OTW, tpos = gw % OTW + 1;
        const __half* wp = p.w + ((size_t)tap * p.Cout + co0) * p.Cin;
        for (int k = 0; k < p.Cin; k += 8) {
            float av[8], l8v[8], a8v[8];
            if (ok) {
                const __half* apk = p.a + operand_unit(bh, WT, CG, twt, k / 8, tpos) * 8;
                load8h(apk, av);
                if (parts == 2) {
                    float lo[8];
                    load8h(apk + a_part, lo);
#pragma unroll
                    for (int e = 0; e < 8; ++e) av[e] += lo[e];
                } else if (parts == 3) {
                    // plane 1: per 16-channel chunk {L8 slab, A8 slab}, 16 bytes per pixel each
                    const uint8_t* p1 = reinterpret_cast<const uint8_t*>(p.a + a_part);
                    const size_t unit = operand_unit(bh, WT, CG, twt, (k / 16) * 2, tpos);
                    const uint2 l8 = *reinterpret_cast<const uint2*>(p1 + unit * 16 + (k & 8));
                    const uint2 a8 = *reinterpret_cast<const uint2*>(p1 + (unit + OPX) * 16 + (k & 8));
#pragma unroll
                    for (int e = 0; e < 8; ++e) {
                        l8v[e] = f8_to_float((uint8_t)(((e < 4 ? l8.x : l8.y) >> (8 * (e & 3))) & 0xff));
                        a8v[e] = f8_to_float((uint8_t)(((e < 4 ? a8.x : a8.y) >> (8 * (e & 3))) & 0xff));
                    }
                }
            } else {
#pragma unroll
                for (int e = 0; e < 8; ++e) av[e] = l8v[e] = a8v[e] = 0.f;
            }
#pragma unroll
            for (int n = 0; n < 4; ++n) {
                float wv[8];
                load8h(wp + (size_t)n * p.Cin + k, wv);
                if (parts == 2) {
                    float lo[8];
                    load8h(wp + w_part + (size_t)n * p.Cin + k, lo);
#pragma unroll
                    for (int e = 0; e < 8; ++e) wv[e] += lo[e];
                }
#pragma unroll
                for (int e = 0; e < 8; ++e) acc[n] = fmaf(av[e], wv[e], acc[n]);
                if (parts == 3) {   // the same two fp8 cross terms the tensor-core path adds
                    float w1[8], w2[8];
                    load8h(wp + w_part + (size_t)n * p.Cin + k, w1);
                    load8h(wp + 2 * w_part + (size_t)n * p.Cin + k, w2);
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[n] = fmaf(l8v[e], w1[e], fmaf(a8v[e], w2[e], acc[n]));
                }
            }
        }
    }
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        float v = 0.f;
        if (valid) {
            const size_t gi = (size_t)pix * p.Cout + co0 + n;
            v = fmaf(acc[n], p.w_inv, p.bias ? p.bias[co0 + n] : 0.f);
            if (p.res) v += p.res[gi];
            v *= p.out_scale;
            p.out[gi] = v;
        }
        if (p.stats) {
            // all 32 lanes of a warp share b when H*W % 32 == 0 (checked on the host)
            const float s1 = warp_sum(v), s2 = warp_sum(v * v);
            if (lane == 0) {
                double* st = p.stats + ((size_t)b * p.Cout + co0 + n) * 2;
                atomicAdd(st, (double)s1);
                atomicAdd(st + 1, (double)s2);
            }
        }
    }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_conv_set_debug(void* dbg_u64) {
    unsigned long long* p = (unsigned long long*)dbg_u64;
    cudaError_t e = cudaMemcpyToSymbol(g_conv_dbg, &p, sizeof(p));
    if (e != cudaSuccess) {
        set_error("conv_set_debug: %s", cudaGetErrorString(e));
        return B200_E_CUDA;
    }
    return B200_OK;
}

extern "C" int b200_conv_set_ablate(int mask) {
#ifndef B200_CONV_ABLATE
    if (mask != 0) {
        set_error("conv_set_ablate: this build has no ablation switches (make EXTRA=-DB200_CONV_ABLATE)");
        return B200_E_ARG;
    }
#endif
    cudaError_t e = cudaMemcpyToSymbol(g_conv_ablate, &mask, sizeof(mask));
    if (e != cudaSuccess) {
        set_error("conv_set_ablate: %s", cudaGetErrorString(e));
        return B200_E_CUDA;
    }
    return B200_OK;
}

extern "C" size_t b200_packed_weight_elems(int Cout, int Cin, int taps, int parts) {
    return (size_t)Cout * Cin * taps * (parts >= 3 ? 2 : parts);   // fp16-sized elements (parts 3: fp16 + 2 x e4m3)
}

static int pack_blocks(size_t total) { return (int)((total + 255) / 256 > 4096 ? 4096 : (total + 255) / 256); }

extern "C" int b200_conv_merged(int bn, int rows, int parts) { return (parts == 2 && 2 * rows * bn <= 256) ? 1 : 0; }

extern "C" int b200_pack_conv_weight(const float* w, void* wpacked, int Cout, int Cin, int taps, int bn, int rows,
                                     int parts, float wscale, void* stream) {
    B200_CHECK_ARG(w && wpacked);
    B200_CHECK_ARG(taps == 9 || taps == 1);
    B200_CHECK_ARG(bn == 64 || bn == 128);
    B200_CHECK_ARG(parts >= 1 && parts <= 4);
    B200_CHECK_ARG(Cout % bn == 0 && Cin % 32 == 0);
    if (rows == 0) {   // column walk (conv_col.cuh): filter rows stacked along N, all weights of the layer in one image
        B200_CHECK_ARG(parts == 3 && taps == 9 && bn == 64 && Cin == 64 && Cout == 64);
        pack_weight_col_kernel<<<pack_blocks((size_t)Cout * Cin * taps), 256, 0, (cudaStream_t)stream>>>(w, (uint8_t*)wpacked, wscale);
        B200_CHECK_LAUNCH();
        return B200_OK;
    }
    if (parts >= 3) {
        const size_t n = (size_t)Cout * Cin * taps;
        pack_weight_f8_kernel<<<pack_blocks(n), 256, 0, (cudaStream_t)stream>>>(w, (__half*)wpacked, Cout, Cin, taps, bn,
                                                                                wscale);
        B200_CHECK_LAUNCH();
        return B200_OK;
    }
    const size_t total = (size_t)Cout * Cin * taps * parts;
    pack_weight_kernel<<<pack_blocks(total), 256, 0, (cudaStream_t)stream>>>(
        w, (__half*)wpacked, Cout, Cin, taps, bn, parts, b200_conv_merged(bn, rows, parts), wscale);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

extern "C" int b200_pack_conv_weight_plain(const float* w, void* w16, int Cout, int Cin, int taps, int parts,
                                           float wscale, void* stream) {
    B200_CHECK_ARG(w && w16);
    B200_CHECK_ARG(taps == 9 || taps == 1);
    B200_CHECK_ARG(parts >= 1 && parts <= 3);
    const size_t total = (size_t)Cout * Cin * taps * parts;
    pack_weight_plain_kernel<<<pack_blocks(total), 256, 0, (cudaStream_t)stream>>>(w, (__half*)w16, Cout, Cin, taps,
                                                                                   parts, wscale);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

// fused front end description (b200_conv_gn_tc); x0 == nullptr: the A operand comes pre-built (b200_conv_tc)
struct GnFront {
    const float* x0 = nullptr;
    const float* x1 = nullptr;
    const double* st0 = nullptr;
    const double* st1 = nullptr;
    const float* gamma = nullptr;
    const float* beta = nullptr;
    const float* ada = nullptr;
    int C0 = 0, C1 = 0, ada_stride = 0, groups = 1, silu = 0;
    float eps = 0.f;
};

static int conv_tc_impl(const void* a, const GnFront& gn, const void* wpacked, const float* bias, const float* res,
                        float out_scale, float w_inv, float* out, double* stats, int B, int H, int W, int Cin, int Cout,
                        int taps, int ring, int bn, int rows, int parts, void* stream, int k_splits = 1) {
    const bool fuse = gn.x0 != nullptr;
    B200_CHECK_ARG((a != nullptr) != fuse);
    B200_CHECK_ARG(wpacked && out);
    B200_CHECK_ARG(B > 0 && H > 0 && W > 0);
    B200_CHECK_ARG(taps == 9 || taps == 1);
    B200_CHECK_ARG(parts >= 1 && parts <= 4);
    B200_CHECK_ARG(W % PIX == 0 && Cin % 32 == 0 && Cin >= 32);
    B200_CHECK_ARG((bn == 64 || bn == 128) && Cout % bn == 0);
    const bool col = fuse && rows == 0;   // column walk (conv_col.cuh): 64 -> 64 channels, 3x3, fp16f8 operands
    if (col) {
        B200_CHECK_ARG(parts == 3 && taps == 9 && bn == 64 && Cin == 64 && Cout == 64 && gn.C1 == 0);
    } else {
        B200_CHECK_ARG((rows == 1 || rows == 2 || rows == 4) && H % rows == 0);
        B200_CHECK_ARG(rows * bn <= 256);   // two TMEM accumulator sets <= 512 columns (merged mode is chosen when 2x fits)
    }
    ConvParams p{};
    p.a = (const __half*)a; p.w = (const __half*)wpacked; p.bias = bias; p.res = res; p.out = out; p.stats = stats;
    p.out_scale = out_scale; p.w_inv = w_inv; p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.ring = ring;
    p.x0 = nullptr; p.x1 = nullptr; p.st0 = nullptr; p.st1 = nullptr;
    p.gn_gamma = nullptr; p.gn_beta = nullptr; p.gn_ada = nullptr;
    p.C0 = p.C1 = p.gn_ada_stride = p.gn_silu = 0; p.gn_groups = 1; p.gn_eps = 0.f;
    p.k_splits = k_splits; p.split_stride = (long long)B * H * W * Cout;
    if (k_splits > 1) {
        B200_CHECK_ARG(!fuse && (Cin / (parts == 1 ? 32 : 16)) % k_splits == 0);
        B200_CHECK_ARG(!bias && !res && !stats);      // the slices hold raw partial sums
    }
    if (fuse) {
        B200_CHECK_ARG(parts == 2 || parts == 3);                               // fp16x3 / fp16f8 operands
        B200_CHECK_ARG(gn.C0 > 0 && gn.C1 >= 0 && gn.C0 + gn.C1 == Cin && Cin <= MAX_CIN);
        B200_CHECK_ARG(gn.C0 % 16 == 0 && gn.C1 % 16 == 0);                     // a 16-channel K chunk never straddles the concat
        B200_CHECK_ARG((gn.C1 > 0) == (gn.x1 != nullptr));
        B200_CHECK_ARG(gn.st0 == nullptr || gn.C1 == 0 || gn.st1 != nullptr);
        B200_CHECK_ARG(gn.groups > 0 && gn.groups <= 32 && Cin % gn.groups == 0);
        B200_CHECK_ARG((gn.gamma == nullptr) == (gn.beta == nullptr));
        p.x0 = gn.x0; p.x1 = gn.x1; p.st0 = gn.st0; p.st1 = gn.st1;
        p.gn_gamma = gn.gamma; p.gn_beta = gn.beta; p.gn_ada = gn.ada;
        p.C0 = gn.C0; p.C1 = gn.C1; p.gn_ada_stride = gn.ada_stride; p.gn_groups = gn.groups; p.gn_silu = gn.silu;
        p.gn_eps = gn.eps;
    }
    cudaStream_t st = (cudaStream_t)stream;
    static int num_sms = 0;
    if (num_sms == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev);
        if (num_sms <= 0) num_sms = 148;
    }
    if (col) return launch_conv_col(p, num_sms, st);
#define B200_CONV_CASE(BN_, R_, NP_)                                                                                  \
    if (!fuse && bn == BN_ && rows == R_ && parts == NP_)                                                             \
        return taps == 9 ? launch_conv<BN_, R_, 9, NP_, false>(p, num_sms, st) : launch_conv<BN_, R_, 1, NP_, false>(p, num_sms, st);
#define B200_FUSE_CASE(BN_, R_, NP_)                                                                                  \
    if (fuse && bn == BN_ && rows == R_ && parts == NP_)                                                              \
        return taps == 9 ? launch_conv<BN_, R_, 9, NP_, true>(p, num_sms, st) : launch_conv<BN_, R_, 1, NP_, true>(p, num_sms, st);
    B200_CONV_CASE(64, 1, 1)
    B200_CONV_CASE(64, 2, 1)
    B200_CONV_CASE(64, 4, 1)
    B200_CONV_CASE(128, 1, 1)
    B200_CONV_CASE(128, 2, 1)
    B200_CONV_CASE(64, 1, 2)
    B200_CONV_CASE(64, 2, 2)
    B200_CONV_CASE(64, 4, 2)
    B200_CONV_CASE(128, 1, 2)
    B200_CONV_CASE(128, 2, 2)
    B200_CONV_CASE(64, 1, 3)
    B200_CONV_CASE(64, 2, 3)
    B200_CONV_CASE(64, 4, 3)
    B200_CONV_CASE(128, 1, 3)
    B200_CONV_CASE(128, 2, 3)
    B200_CONV_CASE(64, 1, 4)
    B200_CONV_CASE(64, 2, 4)
    B200_CONV_CASE(128, 1, 4)
    B200_FUSE_CASE(64, 1, 2)
    B200_FUSE_CASE(64, 2, 2)
    B200_FUSE_CASE(64, 4, 2)
    B200_FUSE_CASE(128, 1, 2)
    B200_FUSE_CASE(128, 2, 2)
    B200_FUSE_CASE(64, 1, 3)
    B200_FUSE_CASE(64, 2, 3)
    B200_FUSE_CASE(64, 4, 3)
    B200_FUSE_CASE(128, 1, 3)
    B200_FUSE_CASE(128, 2, 3)
#undef B200_CONV_CASE
#undef B200_FUSE_CASE
    set_error("conv_tc: unsupported tile bn=%d rows=%d parts=%d fused=%d (need rows*bn <= 256)", bn, rows, parts, (int)fuse);
    return B200_E_ARG;
}

extern "C" int b200_conv_tc(const void* a, const void* wpacked, const float* bias, const float* res, float out_scale,
                            float w_inv, float* out, double* stats, int B, int H, int W, int Cin, int Cout, int taps,
                            int ring, int bn, int rows, int parts, void* stream) {
    B200_CHECK_ARG(a != nullptr);
    return conv_tc_impl(a, GnFront{}, wpacked, bias, res, out_scale, w_inv, out, stats, B, H, W, Cin, Cout, taps, ring, bn,
                        rows, parts, stream);
}

// split-K for the layers with few output tiles and a long reduction (deep levels at small batch: 4x128 C512 at B = 1 is 32 CTAs
// that each walk K = 4608 alone and stream 1.2 MB of weights): `splits` CTAs per tile accumulate disjoint channel slices into
// workspace[splits][B, H*W, Cout] (raw partial sums), conv_splitk_reduce_kernel adds them in slice order and applies the conv
// epilogue (x w_inv + bias + residual, x scale, per-channel statistics).  Same contract as b200_conv_tc.
extern "C" int b200_conv_tc_splitk(const void* a, const void* wpacked, const float* bias, const float* res, float out_scale,
                                   float w_inv, float* out, double* stats, float* workspace, int splits, int B, int H, int W,
                                   int Cin, int Cout, int taps, int ring, int bn, int rows, int parts, void* stream) {
    B200_CHECK_ARG(a != nullptr && workspace != nullptr && out != nullptr);
    B200_CHECK_ARG(splits == 2 || splits == 4 || splits == 8);
    B200_CHECK_ARG(Cout % 4 == 0 && Cout / 4 <= 256 && 256 % (Cout / 4) == 0);
    const int rc = conv_tc_impl(a, GnFront{}, wpacked, nullptr, nullptr, 1.0f, 1.0f, workspace, nullptr, B, H, W, Cin, Cout, taps,
                                ring, bn, rows, parts, stream, splits);
    if (rc != B200_OK) return rc;
    return launch_splitk_reduce(workspace, splits, (size_t)B * H * W * Cout, bias, res, w_inv, out_scale, out, stats, B, H * W, Cout,
                                stream);
}

// GroupNorm(+AdaGN)-apply + SiLU + operand split fused IN FRONT of the conv: replaces the gn_act launch and the operand
// round trip through HBM (reference: efficient_unet.py:104-115 `conv(silu(norm(x)))`, ops.py:176-200, layout_unet_v1.py:229-249)
extern "C" int b200_conv_gn_tc(const float* x0, int C0, const float* x1, int C1, const double* stats0, const double* stats1,
                               const float* gamma, const float* beta, const float* ada, int ada_stride, int groups,
                               float eps, int silu, const void* wpacked, const float* bias, const float* res,
                               float out_scale, float w_inv, float* out, double* stats, int B, int H, int W, int Cout,
                               int taps, int ring, int bn, int rows, int parts, void* stream) {
    B200_CHECK_ARG(x0 != nullptr);
    GnFront gn;
    gn.x0 = x0; gn.x1 = x1; gn.st0 = stats0; gn.st1 = stats1; gn.gamma = gamma; gn.beta = beta; gn.ada = ada;
    gn.C0 = C0; gn.C1 = C1; gn.ada_stride = ada_stride; gn.groups = groups; gn.silu = silu; gn.eps = eps;
    return conv_tc_impl(nullptr, gn, wpacked, bias, res, out_scale, w_inv, out, stats, B, H, W, C0 + C1, Cout, taps, ring,
                        bn, rows, parts, stream);
}

extern "C" int b200_conv_ffma(const void* a, const void* w16, const float* bias, const float* res, float out_scale,
                              float w_inv, float* out, double* stats, int B, int H, int W, int Cin, int Cout, int taps,
                              int ring, int parts, void* stream) {
    B200_CHECK_ARG(a && w16 && out);
    B200_CHECK_ARG(taps == 9 || taps == 1);
    B200_CHECK_ARG(parts >= 1 && parts <= 3);
    B200_CHECK_ARG(Cin % (parts == 3 ? 16 : 8) == 0 && Cout % 32 == 0 && W % OTW == 0);
    B200_CHECK_ARG(!stats || (H * W) % 32 == 0);
    ConvParams p{};
    p.a = (const __half*)a; p.w = (const __half*)w16; p.bias = bias; p.res = res; p.out = out; p.stats = stats;
    p.out_scale = out_scale; p.w_inv = w_inv; p.B = B; p.H = H; p.W = W; p.Cin = Cin; p.Cout = Cout; p.ring = ring;
    const long long npix = (long long)B * H * W;
    dim3 grid((unsigned)((npix + 31) / 32), Cout / 32);
    conv_ffma_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(p, taps, parts);
    B200_CHECK_LAUNCH();
    return B200_OK;
}
