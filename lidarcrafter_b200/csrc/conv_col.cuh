// K1c: "column walk" variant of the fused GroupNorm(+AdaGN)+SiLU -> 3x3 ring conv for the 64 -> 64 channel layers at full
// resolution (efficient_unet.py:104-115 `conv(silu(norm(x)))`, ops.py:32-49,176-200) -- the layers whose ResBlock the
// roofline is reported on.  Included by conv_tc.cu (shares ConvParams, xf_store, the PTX wrappers); entry: b200_conv_gn_tc
// with rows = 0.
//
// Why a second schedule: the tile-walk kernel (conv_tc_kernel<..., FUSE>) stages R + 2 halo rows per R-row tile, so its
// transform warps convert every activation element (R + 2) / R times (2x at R = 2) and the kernel is bound by that
// elementwise instruction stream (profiles/r02_fused_front_ablation.txt).  Here a CTA owns a run of consecutive image rows of
// ONE 128-pixel column:
//   * every input row is converted ONCE into a ring of three row slots in shared memory (whole K = 64 channels per slot:
//     [plane][C/8][130][8], the same tile-major image the separate gn_act kernel writes to HBM) and is read by the MMAs of
//     the three output rows it touches -- the operand never exists in HBM;
//   * the MMAs of an output row run filter row by filter row (dy = 0, 1, 2): the oldest input row is released after the first
//     third of the row's MMAs, so three slots give the transform warps 4/3 of a row time to produce the next input row;
//   * the fp16 halves of the 3x3 weights (72 KB) stay RESIDENT in shared memory for the whole launch; only the e4m3
//     correction halves (72 KB per output row) stream from L2 through a 4-stage ring -- the same L2 traffic per row as the
//     R = 2 tile walk, which re-streams both halves per 2-row tile (an R = 1 walk that streams everything is L2-bound);
//   * accumulators: 4 x 64 TMEM columns (one output row each); the epilogue transposes its 32 x 32 item inside groups of 8
//     lanes with shuffles (no shared-memory staging -- that memory holds the row ring) so that every residual load / output
//     store moves whole 128-byte lines.
// fp16f8 operands only (parts = 3): with fp16x3 the weights are 144 KB of fp16 and cannot stay resident.
#pragma once

#ifndef B200_COL_FENCE_PRODUCER
#define B200_COL_FENCE_PRODUCER 0
#endif

namespace b200 {

struct ColCfg {
    static constexpr int CIN = 64, BN = 64, NCH = 4;
    static constexpr int SLAB = OPX * 16;                // one 8-channel group (16-byte units) of one staged row
    static constexpr int ROW_PLANE = (CIN / 8) * SLAB;   // 16640
    static constexpr int ROW = 2 * ROW_PLANE;            // fp16 hi plane + e4m3 pair plane
    static constexpr int NROW = 3;
    static constexpr int WIMG = 2048;                    // one (chunk, tap) image of one weight plane: [2][64][8] fp16 / [2][64][16] e4m3
    static constexpr int W16 = NCH * 9 * WIMG;           // resident fp16 halves
    static constexpr int B8_STAGE = 6 * WIMG;            // (filter row, chunk pair): 2 chunks x 3 taps
    static constexpr int NB8 = 4;
    static constexpr int NACC = 4;
    static constexpr int THREADS = 640;
    static constexpr int EW = 8;
    static constexpr int OFF_ROWS = 0;
    static constexpr int OFF_W16 = OFF_ROWS + NROW * ROW;
    static constexpr int OFF_B8 = OFF_W16 + W16;
    static constexpr int OFF_STAT = OFF_B8 + NB8 * B8_STAGE;      // [EW][2][32] fp32 partials (flush only)
    static constexpr int OFF_COEF = OFF_STAT + EW * 64 * 4;       // s_a[2][64], s_b[2][64], {mean, rstd}[32]
    static constexpr int OFF_BAR = OFF_COEF + 4 * 64 * 4 + 256;
    static constexpr int SMEM = OFF_BAR + 256;
    static constexpr int TMEM_COLS = NACC * BN;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// whole-row L2 prefetch by the TMA engine: one instruction per 32 KB row (the fp32 activation / residual rows of a 128-pixel
// column are contiguous).  The transform warps keep only ONE row (32 KB per SM) of register loads in flight -- not enough
// to cover HBM latency at 3+ TB/s -- so the producer warp pulls the rows of the next output rows into L2 ahead of them.
__device__ __forceinline__ void l2_prefetch_bulk(const void* ptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
}

// position of a row-tile index t (order: image row fastest, then column, then sample)
struct ColPos { int b, wt, h; };
__device__ __forceinline__ ColPos col_pos(int t, int H, int WT) {
    ColPos r;
    r.h = t % H;
    t /= H;
    r.wt = t % WT;
    r.b = t / WT;
    return r;
}

// 8 x 8 transpose of float4 elements inside every group of 8 lanes (three butterfly steps, 48 shuffles): lane l of a group
// enters with q[j] = channel quad j of ITS pixel and leaves with q[k] = channel quad l of the group's pixel k -- so that the 8
// lanes of a group read / write one whole 128-byte line (32 channels of one pixel) per access.  (A thread-per-pixel epilogue
// straight from the tcgen05.ld layout writes 16-byte pieces of 32 different lines per instruction: measured 9000 cycles per
// 32 x 32 item -- partial-sector writes -- against ~4800 cycles of MMAs per output row.)
__device__ __forceinline__ void group8_transpose(float (&v)[32], int lane) {
#pragma unroll
    for (int m = 4; m >= 1; m >>= 1) {
        const bool upper = (lane & m) != 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (j & m) continue;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float send = upper ? v[4 * j + e] : v[4 * (j | m) + e];
                const float recv = __shfl_xor_sync(0xffffffffu, send, m);
                if (upper) v[4 * j + e] = recv;
                else v[4 * (j | m) + e] = recv;
            }
        }
    }
}

template <bool SILU>
__device__ __forceinline__ void col_xform_warps(const ConvParams& p, uint32_t sbase, uint32_t bar0, const float* s_coef,
                                                int tw, int lane, int t_lo, int t_hi, int b_lo, int ablate) {
    using C = ColCfg;
    (void)ablate;
    const int WT = p.W / PIX;
    const int g = lane >> 4, p8 = (lane >> 1) & 7, q = lane & 1;
    const int co = g * 8 + q * 4;                           // this lane's 4 channels inside a 16-channel chunk
    const int px0 = tw * 8 + p8;                            // pixel inside the tile (first group; second: + 64)
    // byte offsets of this lane inside a row slot: unit (e, c) adds e * 64 * 16 + c * 2 * SLAB
    const uint32_t so_hi = sbase + C::OFF_ROWS + g * C::SLAB + (1 + px0) * 16 + q * 8;
    const uint32_t so_p1 = sbase + C::OFF_ROWS + C::ROW_PLANE + (1 + px0) * 16 + co;
    // ring-halo duty (one warp per row, in turns): lane -> side = lane / 16, channel quad cq = lane % 16
    const int h_side = lane >> 4, h_cq = lane & 15;
    const uint32_t h_pos = (h_side ? OPX - 1 : 0) * 16;
    const uint32_t ho_hi = sbase + C::OFF_ROWS + (h_cq >> 1) * C::SLAB + h_pos + (h_cq & 1) * 8;
    const uint32_t ho_p1 = sbase + C::OFF_ROWS + C::ROW_PLANE + (h_cq >> 2) * 2 * C::SLAB + h_pos + (h_cq & 3) * 4;

    // input-row stream: for every output row-tile t of [t_lo, t_hi): rows h-1, h, h+1 if t starts a run (first tile of the
    // CTA or of a column), else only h+1
    auto is_first = [&](int t) { return t == t_lo || (t % p.H) == 0; };
    struct InRow { int b, wt, h; bool img; };
    auto decode = [&](int t, int k) {
        const ColPos cp = col_pos(t, p.H, WT);
        InRow r;
        r.b = cp.b; r.wt = cp.wt; r.h = cp.h - 1 + k;
        r.img = (unsigned)r.h < (unsigned)p.H;
        return r;
    };
    float4 ring[8], hring = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 8; ++u) ring[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* l_ptr = nullptr;      // this lane's float4 of unit (e = 0, c = 0) of the row being loaded
    const float* l_hptr = nullptr;
    bool l_on = false, l_h = false;
    auto load_ctx = [&](const InRow& r, bool on, uint32_t ir_of) {
        l_on = on && r.img;
        l_ptr = p.x0 + ((size_t)(r.b * p.H + r.h) * p.W + r.wt * PIX + px0) * C::CIN + co;
        int ww = r.wt * PIX + (h_side ? PIX : -1);
        bool ok = l_on && (int)(ir_of & 7u) == tw;
        if (ww < 0) { ww += p.W; ok = ok && p.ring; }
        else if (ww >= p.W) { ww -= p.W; ok = ok && p.ring; }
        l_h = ok;
        l_hptr = p.x0 + ((size_t)(r.b * p.H + r.h) * p.W + ww) * C::CIN + h_cq * 4;
    };
    auto load_unit = [&](int u) {      // u = e * 4 + c (compile-time)
        if (l_on && !ABL(64)) ring[u] = ldg_stream_f4(l_ptr + (u >> 2) * 64 * C::CIN + (u & 3) * 16);
    };

    unsigned long long* dbg = (tw == 0 && lane == 0) ? g_conv_dbg : nullptr;
    unsigned long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_start = dbg ? clock64() : 0;
    int t = t_lo, k = 0;
    uint32_t ir = 0;
    {
        const InRow r0 = decode(t, k);
        load_ctx(r0, true, 0);
#pragma unroll
        for (int u = 0; u < 8; ++u) load_unit(u);
        if (l_h) hring = ldg_stream_f4(l_hptr);
    }
    mbar_wait_quiet(bar0 + 184u, 0);     // COEF_FULL
    while (t < t_hi) {
        const InRow cur = decode(t, k);
        int t2 = t, k2 = k + 1;
        if (k2 > 2) { t2 = t + 1; k2 = (t2 < t_hi && is_first(t2)) ? 0 : 2; }
        const bool have_next = t2 < t_hi;
        const InRow nx = decode(have_next ? t2 : t, have_next ? k2 : k);
        const uint32_t slot = ir % C::NROW;
        {
            DBG_T0();
            mbar_wait_quiet(bar0 + 24u + 8u * slot, ((ir / C::NROW) & 1) ^ 1);      // EMPTY_ROW(slot)
            DBG_ACC(6);
        }
        const uint32_t st = slot * C::ROW;
        const bool halo_duty = (int)(ir & 7u) == tw;
        if (ABL(128)) {
            load_ctx(nx, have_next, ir + 1);
        } else if (cur.img) {
            const float* sa = s_coef + (cur.b - b_lo) * 64;
            const float* sb = sa + 128;
            load_ctx(nx, have_next, ir + 1);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                const float4 a4 = *reinterpret_cast<const float4*>(sa + c * 16 + co);
                const float4 b4 = *reinterpret_cast<const float4*>(sb + c * 16 + co);
                const float ca[4] = {a4.x, a4.y, a4.z, a4.w}, cb[4] = {b4.x, b4.y, b4.z, b4.w};
                float4 x[2];
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    x[e] = ring[e * 4 + c];
                    load_unit(e * 4 + c);
                }
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    const uint32_t off = st + e * 64 * 16 + c * 2 * C::SLAB;
                    xf_store<3, SILU, false>(x[e], ca, cb, true, so_hi + off, so_p1 + off, so_p1 + off + C::SLAB, ablate);
                }
            }
            {
                const float4 x4 = hring;
                if (l_h) hring = ldg_stream_f4(l_hptr);
                if (halo_duty) {
                    const int ww = cur.wt * PIX + (h_side ? PIX : -1);
                    const bool valid = p.ring || (ww >= 0 && ww < p.W);
                    const float4 a4 = *reinterpret_cast<const float4*>(sa + h_cq * 4);
                    const float4 b4 = *reinterpret_cast<const float4*>(sb + h_cq * 4);
                    const float ha[4] = {a4.x, a4.y, a4.z, a4.w}, hb[4] = {b4.x, b4.y, b4.z, b4.w};
                    xf_store<3, SILU, true>(x4, ha, hb, valid, ho_hi + st, ho_p1 + st, ho_p1 + st + C::SLAB);
                }
            }
        } else {
            // row outside the image: exact zeros (not act(b)); the register ring is free, request the whole next row
            load_ctx(nx, have_next, ir + 1);
#pragma unroll
            for (int u = 0; u < 8; ++u) load_unit(u);
            if (l_h) hring = ldg_stream_f4(l_hptr);
            const uint32_t base = sbase + C::OFF_ROWS + st;
            for (int i = tw * 32 + lane; i < C::ROW / 16; i += XF_WARPS * 32)
                asm volatile("st.shared.v4.b32 [%0], {%1, %1, %1, %1};" ::"r"(base + i * 16), "r"(0u) : "memory");
        }
        // The generic-proxy -> async-proxy fence for these stores is executed by the CONSUMER (the MMA issuer warp, after its
        // acquire of FULL_ROW): here ptxas lowers fence.proxy.async to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, and the MEMBAR also
        // waits for this warp's prefetch loads of the NEXT row -- one exposed HBM round trip per row (5450 -> ... cycles per
        // row).  The arrive below is a release at CTA scope: the stores are performed before the phase completes.
#if B200_COL_FENCE_PRODUCER
        fence_proxy_async();
#endif
        __syncwarp();
        if (lane == 0) mbar_arrive(bar0 + 8u * slot);                           // FULL_ROW(slot)
        ++ir;
        t = t2;
        k = k2;
    }
    if (dbg) {
        dbg[blockIdx.x * 8 + 6] = dbg_acc[6];
        dbg[blockIdx.x * 8 + 7] = clock64() - t_start;
    }
}

__global__ void __launch_bounds__(ColCfg::THREADS, 1) conv_col_kernel(const ConvParams p) {
    using C = ColCfg;
    extern __shared__ __align__(128) uint8_t smem[];
    const uint32_t sbase = smem_u32(smem);
    const uint32_t bar0 = sbase + C::OFF_BAR;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + C::OFF_BAR + 240);
#define CFULL_ROW(s) (bar0 + 8u * (s))
#define CEMPTY_ROW(s) (bar0 + 24u + 8u * (s))
#define CFULL_B8(s) (bar0 + 48u + 8u * (s))
#define CEMPTY_B8(s) (bar0 + 80u + 8u * (s))
#define CACC_FULL(s) (bar0 + 112u + 8u * (s))
#define CACC_EMPTY(s) (bar0 + 144u + 8u * (s))
#define CW_FULL (bar0 + 176u)
#define CCOEF_FULL (bar0 + 184u)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int WT = p.W / PIX;
    const int per_cta = (p.n_tiles + gridDim.x - 1) / gridDim.x;
    const int t_lo = blockIdx.x * per_cta;
    const int t_hi = min(t_lo + per_cta, p.n_tiles);
    const int per_sample = WT * p.H;
    const int b_lo = t_lo / per_sample, b_hi = (max(t_hi, t_lo + 1) - 1) / per_sample;   // <= b_lo + 1 (checked on the host)

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::NROW; ++s) {
            mbar_init(CFULL_ROW(s), XF_WARPS);
            mbar_init(CEMPTY_ROW(s), 2);           // one commit from each MMA issuer warp
        }
        for (int s = 0; s < C::NB8; ++s) {
            mbar_init(CFULL_B8(s), 1);
            mbar_init(CEMPTY_B8(s), 1);
        }
        for (int s = 0; s < C::NACC; ++s) {
            mbar_init(CACC_FULL(s), 2);
            mbar_init(CACC_EMPTY(s), C::EW * 32);
        }
        mbar_init(CW_FULL, 1);
        mbar_init(CCOEF_FULL, 1);
        fence_barrier_init();
    }
    pdl_launch_dependents();
    if (warp == 5) tmem_alloc(smem_u32(tmem_slot), C::TMEM_COLS);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    pdl_wait();

    const int ablate = g_conv_ablate;
    (void)ablate;
    const int wg = warp >> 2;
    if (wg >= 3) {
        reg_dealloc<88>();
        if (t_lo < t_hi) {
            const float* s_coef = reinterpret_cast<const float*>(smem + C::OFF_COEF);
            if (p.gn_silu) col_xform_warps<true>(p, sbase, bar0, s_coef, warp - 12, lane, t_lo, t_hi, b_lo, ablate);
            else col_xform_warps<false>(p, sbase, bar0, s_coef, warp - 12, lane, t_lo, t_hi, b_lo, ablate);
        }
    } else if (wg == 1) {
        reg_dealloc<56>();
        if (warp == 4) {
            // ------------------------------ coefficient warp: y = x * s_a[c] + s_b[c] for the (<= 2) samples of this CTA ------------------------------
            if (t_lo < t_hi) {
                float* s_a = reinterpret_cast<float*>(smem + C::OFF_COEF);      // [2][64]
                float* s_b = s_a + 128;                                         // [2][64]
                float* s_mr = s_a + 256;                                        // {mean, rstd}[32]
                for (int b = b_lo; b <= b_hi; ++b) {
                    float* sa = s_a + (b - b_lo) * 64;
                    float* sb = s_b + (b - b_lo) * 64;
                    if (p.st0 != nullptr) {
                        const int cpg = C::CIN / p.gn_groups;
                        if (lane < p.gn_groups) {
                            double su = 0.0, ss = 0.0;
                            for (int i = 0; i < cpg; ++i) {
                                const double2 st = *reinterpret_cast<const double2*>(p.st0 + ((size_t)b * C::CIN + lane * cpg + i) * 2);
                                su += st.x;
                                ss += st.y;
                            }
                            const double n = (double)p.H * p.W * cpg;
                            const double mean = su / n;
                            double var = ss / n - mean * mean;
                            if (var < 0.0) var = 0.0;
                            s_mr[2 * lane] = (float)mean;
                            s_mr[2 * lane + 1] = (float)(1.0 / sqrt(var + (double)p.gn_eps));
                        }
                        __syncwarp();
                        for (int c = lane; c < C::CIN; c += 32) {
                            const int gi = c / cpg;
                            float a = s_mr[2 * gi + 1], bb = -s_mr[2 * gi] * s_mr[2 * gi + 1];
                            float ga = 1.f, be = 0.f, sc = 1.f, sh = 0.f;
                            if (p.gn_gamma) { ga = p.gn_gamma[c]; be = p.gn_beta[c]; }
                            if (p.gn_ada) {
                                sc = 1.f + p.gn_ada[(size_t)b * p.gn_ada_stride + c];
                                sh = p.gn_ada[(size_t)b * p.gn_ada_stride + C::CIN + c];
                            }
                            a *= ga; bb = bb * ga + be;
                            a *= sc; bb = bb * sc + sh;
                            sa[c] = a;
                            sb[c] = bb;
                        }
                        __syncwarp();
                    } else {
                        for (int c = lane; c < C::CIN; c += 32) { sa[c] = 1.f; sb[c] = 0.f; }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(CCOEF_FULL);
            }
        } else if (warp == 6) {
            // ------------------------------ weight producer: resident fp16 halves once, e4m3 halves per (row, dy, chunk pair) ------------------------------
            if (lane == 0 && t_lo < t_hi) {
                const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w);
                mbar_expect_tx(CW_FULL, C::W16);
                for (int i = 0; i < C::NCH * 9; ++i)
                    bulk_copy_g2s(sbase + C::OFF_W16 + i * C::WIMG, wsrc + (size_t)i * 2 * C::WIMG, C::WIMG, CW_FULL);
                // L2 prefetch distance in output rows: input row h + 1 and the residual row of tile t + PF
                constexpr int PF = 4;
                constexpr uint32_t ROWB = PIX * C::CIN * 4;
                auto prefetch_rows = [&](int tp) {
                    if (tp >= t_hi) return;
                    const ColPos cp = col_pos(tp, p.H, WT);
                    const size_t base = ((size_t)(cp.b * p.H + cp.h) * p.W + cp.wt * PIX) * C::CIN;
                    const bool first = tp == t_lo || cp.h == 0;
                    if (first) {
                        if (cp.h > 0) l2_prefetch_bulk(p.x0 + base - (size_t)p.W * C::CIN, ROWB);
                        l2_prefetch_bulk(p.x0 + base, ROWB);
                    }
                    if (cp.h + 1 < p.H) l2_prefetch_bulk(p.x0 + base + (size_t)p.W * C::CIN, ROWB);
                    if (p.res) l2_prefetch_bulk(p.res + base, ROWB);
                };
                for (int i = 0; i < PF; ++i) prefetch_rows(t_lo + i);
                uint32_t g = 0;
                for (int t = t_lo; t < t_hi; ++t) {
                    prefetch_rows(t + PF);
                    for (int u = 0; u < 6; ++u, ++g) {
                        const int dy = u >> 1, half = u & 1;
                        const uint32_t s = g % C::NB8;
                        mbar_wait(CEMPTY_B8(s), ((g / C::NB8) & 1) ^ 1);
                        mbar_expect_tx(CFULL_B8(s), C::B8_STAGE);
#pragma unroll
                        for (int j = 0; j < 6; ++j) {
                            const int c = 2 * half + j / 3, tap = dy * 3 + j % 3;
                            bulk_copy_g2s(sbase + C::OFF_B8 + s * C::B8_STAGE + j * C::WIMG,
                                          wsrc + (size_t)(c * 9 + tap) * 2 * C::WIMG + C::WIMG, C::WIMG, CFULL_B8(s));
                        }
                    }
                }
            }
        } else {
            // ------------------------------ MMA issuers (warps 5 / 7): issue unit = (output row, filter row dy, chunk pair) ------------------------------
            // warp 5 takes chunk pair 0, warp 7 chunk pair 1 of every (row, dy): they alternate on the tensor pipe (named
            // barriers 2 / 3) exactly like the tile-walk kernel's issuers; each commits what ITS MMAs read.
            if (t_lo < t_hi) {
                const uint32_t par = warp == 7 ? 1u : 0u;
                constexpr uint32_t idesc = make_idesc_f16(128, C::BN);
                unsigned long long* dbg = (lane == 0 && par == 0) ? g_conv_dbg : nullptr;
                unsigned long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                const long long t_start = dbg ? clock64() : 0;
                mbar_wait(CW_FULL, 0);
                uint32_t it = 0, base = 0, ir_count = 0;
                for (int t = t_lo; t < t_hi; ++t, ++it) {
                    const int h = t % p.H;
                    const bool first = t == t_lo || h == 0, last = t == t_hi - 1 || h == p.H - 1;
                    if (first) { base = ir_count; ir_count += 3; }
                    else { base += 1; ir_count += 1; }
                    const uint32_t buf = it % C::NACC;
                    const uint32_t acc = tmem_base + buf * C::BN;
#pragma unroll 1
                    for (int dy = 0; dy < 3; ++dy) {
                        const uint32_t g = 6u * it + 2u * dy + par;
                        if (dy == 0) {
                            DBG_T0();
                            mbar_wait(CACC_EMPTY(buf), ((it / C::NACC) & 1) ^ 1);
                            DBG_ACC(3);
                        }
                        const uint32_t irow = base + dy, slot = irow % C::NROW;
                        {
                            DBG_T0();
                            mbar_wait(CFULL_ROW(slot), (irow / C::NROW) & 1);
                            DBG_ACC(1);
                        }
#if !B200_COL_FENCE_PRODUCER
                        fence_proxy_async();     // the transform warps' st.shared (acquired above) -> async-proxy operand reads
#endif
                        const uint32_t sb = g % C::NB8;
                        {
                            DBG_T0();
                            mbar_wait(CFULL_B8(sb), (g / C::NB8) & 1);
                            DBG_ACC(2);
                        }
                        // descriptor bases of this unit: A = row slot (hi plane | pair plane), chunk 2 par; B = resident fp16 image of
                        // (chunk, tap dy * 3), streamed e4m3 image of the stage
                        const uint32_t a16 = desc_lo(sbase + C::OFF_ROWS + slot * C::ROW + (2 * par) * 2 * C::SLAB, C::SLAB);
                        const uint32_t a8 = a16 + (C::ROW_PLANE >> 4);
                        const uint32_t b16 = desc_lo(sbase + C::OFF_W16 + ((2 * par) * 9 + dy * 3) * C::WIMG, C::BN * 16);
                        const uint32_t b8 = desc_lo(sbase + C::OFF_B8 + sb * C::B8_STAGE, C::BN * 16);
                        if (g != 0) {
                            DBG_T0();
                            named_bar_sync(2 + par, 64);
                            DBG_ACC(4);
                        }
                        tc_fence_after();
                        if (elect_one()) {
                            if (!ABL(8))
#pragma unroll
                            for (int cc = 0; cc < 2; ++cc) {
#pragma unroll
                                for (int dx = 0; dx < 3; ++dx) {
                                    const uint32_t accum = (dy | (int)par | cc | dx) != 0 ? 1u : 0u;
                                    const uint32_t ao = (uint32_t)((cc * 2 * C::SLAB + dx * 16) >> 4);
                                    tc_mma_f16_lh(acc, a16 + ao, b16 + (uint32_t)(((cc * 9 + dx) * C::WIMG) >> 4), idesc, accum);
                                    tc_mma_f8_lh(acc, a8 + ao, b8 + (uint32_t)(((cc * 3 + dx) * C::WIMG) >> 4), idesc, 1u);
                                }
                            }
                            tc_commit(CEMPTY_B8(sb));
                            if (dy == 0 || last) tc_commit(CEMPTY_ROW(slot));
                            if (dy == 2) tc_commit(CACC_FULL(buf));
                        }
                        __syncwarp();
                        tc_fence_before();
                        named_bar_arrive(2 + (par ^ 1u), 64);
                    }
                }
                if (dbg) {
                    dbg[blockIdx.x * 8 + 0] = clock64() - t_start;
                    dbg[blockIdx.x * 8 + 1] = dbg_acc[1];
                    dbg[blockIdx.x * 8 + 2] = dbg_acc[2];
                    dbg[blockIdx.x * 8 + 3] = dbg_acc[3];
                }
            }
        }
    } else {
        reg_alloc<120>();
        // ------------------------------ epilogue: warps 0-3 and 8-11; one item (32 pixels x 32 channels) per warp and output row ------------------------------
        if (t_lo < t_hi) {
            const int ew = warp < 4 ? warp : warp - 4;
            const int quarter = warp & 3, slice = ew >> 2;
            const float scale = p.out_scale, winv = p.w_inv;
            unsigned long long* dbg = (threadIdx.x == 0) ? g_conv_dbg : nullptr;
            unsigned long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            const long long t_start = dbg ? clock64() : 0;
            // lane = (pixel group P = lane / 8, l = lane % 8): after the transpose it owns channel quad l (4 channels) of the pixels
            // 8 P .. 8 P + 7 of this warp's 32-pixel quarter
            const int P8 = lane >> 3, l8 = lane & 7;
            float acc1[4] = {0.f, 0.f, 0.f, 0.f}, acc2[4] = {0.f, 0.f, 0.f, 0.f};   // running per-channel sum / sum of squares
            float* sst = reinterpret_cast<float*>(smem + C::OFF_STAT);
            const float4 bi = p.bias ? *reinterpret_cast<const float4*>(p.bias + slice * 32 + 4 * l8) : make_float4(0, 0, 0, 0);
            int cur_b = -1;
            auto flush_stats = [&]() {
                if (lane < 8) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        sst[ew * 64 + 4 * l8 + e] = acc1[e];
                        sst[ew * 64 + 32 + 4 * l8 + e] = acc2[e];
                    }
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) acc1[e] = acc2[e] = 0.f;
                named_bar_sync(1, C::EW * 32);
                if (ew < 2) {                    // warp ew = 0 / 1 reduces slice ew: lanes over its 32 channels
                    const int sl = ew;
                    float a = 0.f, q2 = 0.f;
#pragma unroll
                    for (int w = 0; w < 4; ++w) {            // the four lane quarters of slice sl: epilogue warps sl * 4 + w
                        a += sst[(sl * 4 + w) * 64 + lane];
                        q2 += sst[(sl * 4 + w) * 64 + 32 + lane];
                    }
                    double* st = p.stats + ((size_t)cur_b * C::BN + sl * 32 + lane) * 2;
                    atomicAdd(st, (double)a);
                    atomicAdd(st + 1, (double)q2);
                }
                named_bar_sync(1, C::EW * 32);
            };
            uint32_t it = 0;
            for (int t = t_lo; t < t_hi; ++t, ++it) {
                const ColPos cp = col_pos(t, p.H, WT);
                if (p.stats && cp.b != cur_b) {
                    if (cur_b >= 0) flush_stats();
                    cur_b = cp.b;
                }
                const uint32_t buf = it % C::NACC;
                // element offset of (pixel 8 P + k, channel quad l): k adds BN
                const size_t gi = ((size_t)(cp.b * p.H + cp.h) * p.W + cp.wt * PIX + quarter * 32 + P8 * 8) * C::BN + slice * 32 + 4 * l8;
                float4 rv[8];
                const bool do_res = p.res && !ABL(4);
                if (do_res) {
#pragma unroll
                    for (int k = 0; k < 8; ++k) rv[k] = *reinterpret_cast<const float4*>(p.res + gi + (size_t)k * C::BN);
                }
                {
                    DBG_T0();
                    mbar_wait(CACC_FULL(buf), (it / C::NACC) & 1);
                    DBG_ACC(5);
                }
                tc_fence_after();
                float v[32];
                tmem_ld_32x32(tmem_base + buf * C::BN + slice * 32 + ((uint32_t)(quarter * 32) << 16), v);
                tc_fence_before();
                mbar_arrive(CACC_EMPTY(buf));
                group8_transpose(v, lane);
                float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    float4 tv;
                    tv.x = fmaf(v[4 * k], winv, bi.x); tv.y = fmaf(v[4 * k + 1], winv, bi.y);
                    tv.z = fmaf(v[4 * k + 2], winv, bi.z); tv.w = fmaf(v[4 * k + 3], winv, bi.w);
                    if (do_res) { tv.x += rv[k].x; tv.y += rv[k].y; tv.z += rv[k].z; tv.w += rv[k].w; }
                    tv.x *= scale; tv.y *= scale; tv.z *= scale; tv.w *= scale;
                    if (!ABL(4)) *reinterpret_cast<float4*>(p.out + gi + (size_t)k * C::BN) = tv;
                    s1[0] += tv.x; s1[1] += tv.y; s1[2] += tv.z; s1[3] += tv.w;
                    s2[0] += tv.x * tv.x; s2[1] += tv.y * tv.y; s2[2] += tv.z * tv.z; s2[3] += tv.w * tv.w;
                }
                if (p.stats) {
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], 8);
                        s1[e] += __shfl_xor_sync(0xffffffffu, s1[e], 16);
                        s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], 8);
                        s2[e] += __shfl_xor_sync(0xffffffffu, s2[e], 16);
                        acc1[e] += s1[e];
                        acc2[e] += s2[e];
                    }
                }
            }
            if (p.stats && cur_b >= 0) flush_stats();
            if (dbg) {
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
#undef CFULL_ROW
#undef CEMPTY_ROW
#undef CFULL_B8
#undef CEMPTY_B8
#undef CACC_FULL
#undef CACC_EMPTY
#undef CW_FULL
#undef CCOEF_FULL
}

static int launch_conv_col(ConvParams p, int num_sms, cudaStream_t st) {
    using C = ColCfg;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_col_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, C::SMEM);
        if (e != cudaSuccess) {
            set_error("conv_col: cudaFuncSetAttribute(%d B smem) failed: %s", C::SMEM, cudaGetErrorString(e));
            return B200_E_CUDA;
        }
        attr_set = true;
    }
    p.n_tiles = (p.W / PIX) * p.H * p.B;        // 128-pixel row tiles
    int grid = p.n_tiles < num_sms ? p.n_tiles : num_sms;
    const int per = (p.n_tiles + grid - 1) / grid;
    grid = (p.n_tiles + per - 1) / per;
    if (per > (p.W / PIX) * p.H) {              // a CTA's run would span more than two samples (coefficient slots)
        set_error("conv_col: batch too large for the column walk (B = %d)", p.B);
        return B200_E_ARG;
    }
    launch_pdl_if(pdl_enabled_conv(), conv_col_kernel, dim3(grid), dim3(C::THREADS), (size_t)C::SMEM, st, p);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

}  // namespace b200
