// K1c: "column walk" variant of the fused GroupNorm(+AdaGN)+SiLU -> 3x3 ring conv for the 64 -> 64 channel layers at full
// resolution (efficient_unet.py:104-115 `conv(silu(norm(x)))`, ops.py:32-49,176-200) -- the layers whose ResBlock the
// roofline is reported on.  Included by conv_tc.cu (shares ConvParams, the PTX wrappers); entry: b200_conv_gn_tc with
// rows = 0 (weights packed by b200_pack_conv_weight with rows = 0).
//
// Why a second schedule.  Measured on the tile-walk kernel (conv_tc_kernel<64, 2, 9, 3, FUSE>) and on the first versions of
// this one (profiles/r02_fused_front_ablation.txt, profiles/r02_col_walk.txt):
//   * the tile walk stages R + 2 halo rows per R-row tile: its transform warps convert every element (R + 2) / R times;
//   * an M = 128 x N = 64 MMA reads 4 KB of A + 2 KB of B from shared memory in 48 cycles = the whole 128 B/clk of the SM's
//     shared-memory / L1 data path, which the transform warps' st.shared, every global load return and the epilogue also need:
//     with 72 such MMAs per output row everything else starves (a transform body that costs 500 cycles per row alone took
//     5300 next to the MMAs);
//   * issue units of 12 MMAs (~580 cycles) are shorter than the polls + hand-over between the two issuer warps (~650).
// Hence an INPUT-ROW-STATIONARY schedule.  A CTA owns a run of consecutive image rows of ONE 128-pixel column:
//   * every input row is converted ONCE (GroupNorm-apply, SiLU, fp16 | e4m3 split) into shared memory -- two half-row slots
//     of 32 channels each, four half slots in the ring -- in the same tile-major image the separate gn_act kernel writes to
//     HBM; the operand never exists in HBM;
//   * input row r contributes to the output rows r + 1, r, r - 1 through the filter rows dy = 0, 1, 2.  The weight image stacks
//     the three filter rows along N ([dy][cout] = 192 rows per (chunk, dx)), and the accumulators of consecutive output rows sit
//     in a ring of 8 x 64 TMEM columns in DESCENDING order, so ONE M = 128 x N = 192 MMA per (chunk, dx, operand kind)
//     adds a row's contribution to all three: 24 MMAs of 96 cycles per row instead of 72 of 48, and 240 KB instead of 432 KB
//     of shared-memory operand reads per row.  Run / column borders use N = 64 / 128 sub-ranges of the same image, the ring
//     wrap-around splits an MMA in two; the epilogue hands every accumulator back zeroed (all MMAs accumulate);
//   * ALL weights (fp16 halves + e4m3 correction halves, 144 KB) are resident in shared memory for the whole launch: no
//     weight streaming from L2 (the R = 2 tile walk re-reads them per 2-row tile: 150 MB per launch);
//   * the epilogue transposes its 32 x 32 item inside groups of 8 lanes with shuffles (no shared-memory staging -- that
//     memory holds the weights) so that every residual load / output store moves whole 128-byte lines.
// fp16f8 operands only (parts = 3): with fp16x3 the weights are 288 KB.
#pragma once

#ifndef B200_COL_FENCE_PRODUCER
#define B200_COL_FENCE_PRODUCER 0
#endif

namespace b200 {

struct ColCfg {
    static constexpr int CIN = 64, BN = 64, NCH = 4;
    static constexpr int SLAB = OPX * 16;                // one 8-channel group (16-byte units) of one staged row
    static constexpr int HALF_PLANE = 4 * SLAB;          // 32 channels of one plane
    static constexpr int HALF = 2 * HALF_PLANE;          // half-row slot: fp16 hi plane + e4m3 pair plane of 32 channels
    static constexpr int NHALF = 4;
    static constexpr int WROWS = 3 * BN;                 // stacked filter rows: N = 192
    static constexpr int WIMG = 2 * WROWS * 16;          // one (chunk, dx, plane) image: [2 k-groups][192][16 B]
    static constexpr int WALL = NCH * 3 * 2 * WIMG;      // 147456
    static constexpr int NACC = 8;
    static constexpr int THREADS = 640;
    static constexpr int EW = 8;
    static constexpr int OFF_ROWS = 0;
    static constexpr int OFF_W = OFF_ROWS + NHALF * HALF;
    static constexpr int OFF_STAT = OFF_W + WALL;                 // [EW][2][32] fp32 partials (flush only)
    static constexpr int OFF_COEF = OFF_STAT + EW * 64 * 4;       // s_a[2][64], s_b[2][64], {mean, rstd}[32]
    static constexpr int OFF_BAR = OFF_COEF + 4 * 64 * 4 + 256;
    static constexpr int SMEM = OFF_BAR + 256;
    static constexpr int TMEM_COLS = NACC * BN;
    static_assert(SMEM <= 227 * 1024, "shared memory budget");
};

// whole-row L2 prefetch by the TMA engine: one instruction per 32 KB row (the fp32 activation / residual rows of a 128-pixel
// column are contiguous).  The transform warps keep only ONE row (32 KB per SM) of register loads in flight.
__device__ __forceinline__ void l2_prefetch_bulk(const void* ptr, uint32_t bytes) {
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(ptr), "r"(bytes) : "memory");
}

__device__ __forceinline__ void tmem_st_zero_32x32(uint32_t taddr) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
        "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(0u)
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
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

// The INPUT rows of a CTA's run [t_lo, t_hi) of output row tiles.  The run is cut into segments (one column each: output
// rows h_a .. h_b); a segment reads the input rows max(h_a - 1, 0) .. min(h_b + 1, H - 1) -- rows outside the image
// contribute nothing and are never staged.  it_a: running index of the segment's first output row inside the CTA.
struct ColRows {
    int t_hi, H, WT;
    int t, h_a, h_b, rho, rho_b, it_a, b, wt;      // (b, wt): the segment's column -- kept incrementally (no division per row)
    __device__ __forceinline__ void seg() {
        const int n = min(H - h_a, t_hi - t);
        h_b = h_a + n - 1;
        rho = max(h_a - 1, 0);
        rho_b = min(h_b + 1, H - 1);
    }
    __device__ __forceinline__ void init(int t_lo, int t_hi_, int H_, int WT_) {
        t_hi = t_hi_; H = H_; WT = WT_; t = t_lo; it_a = 0;
        const ColPos cp = col_pos(t_lo, H, WT);
        b = cp.b; wt = cp.wt; h_a = cp.h;
        h_b = rho = rho_b = 0;
        if (t < t_hi) seg();
    }
    __device__ __forceinline__ bool valid() const { return t < t_hi; }
    __device__ __forceinline__ void next() {
        if (rho < rho_b) { ++rho; return; }
        const int n = h_b - h_a + 1;
        t += n;
        it_a += n;
        if (t < t_hi) {             // a segment that does not end the run ends a column: the next one starts the next column
            h_a = 0;
            if (++wt == WT) { wt = 0; ++b; }
            seg();
        }
    }
};

// Waiting without eating issue slots: mbarrier.try_wait returns after a short, system-defined time however long the hint, and a
// tight retry loop costs ~8 instructions per ~60 cycles -- measured (ncu source page) at half of all instructions the 8 epilogue
// warps execute, on an SM whose schedulers are 80 % busy.  Warps with slack (epilogue: 8 accumulators deep) sleep between polls.
template <int NS>
__device__ __forceinline__ void mbar_wait_sleepy(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        __nanosleep(NS);
        if (++spins > (1u << 22)) __trap();
    }
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

// GroupNorm-apply (+SiLU) + fp16 | e4m3 split of one lane's 4 channels: the same expressions as gn_act_kernel / xf_store
// (bit-identical operands): w0 w1 = fp16 hi, w2 = e4m3(lo * 2^11) x 4, w3 = e4m3(y) x 4
template <bool SILU>
__device__ __forceinline__ void col_convert(const float4 x4, const float (&ca)[4], const float (&cb)[4], bool valid, uint32_t& w0,
                                            uint32_t& w1, uint32_t& w2, uint32_t& w3) {
    float y[4] = {fmaf(x4.x, ca[0], cb[0]), fmaf(x4.y, ca[1], cb[1]), fmaf(x4.z, ca[2], cb[2]), fmaf(x4.w, ca[3], cb[3])};
    if (SILU) {
#pragma unroll
        for (int e = 0; e < 4; ++e) y[e] = silu_f(y[e]);
    }
    const __half2 h01 = __floats2half2_rn(y[0], y[1]), h23 = __floats2half2_rn(y[2], y[3]);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const float lo[4] = {y[0] - f01.x, y[1] - f01.y, y[2] - f23.x, y[3] - f23.y};
    w0 = *reinterpret_cast<const uint32_t*>(&h01);
    w1 = *reinterpret_cast<const uint32_t*>(&h23);
    w2 = f8x4(lo[0] * F8_LO_SCALE, lo[1] * F8_LO_SCALE, lo[2] * F8_LO_SCALE, lo[3] * F8_LO_SCALE);
    w3 = f8x4(y[0], y[1], y[2], y[3]);
    if (!valid) w0 = w1 = w2 = w3 = 0u;     // zero padding (non-ring edges) is exact: zeros, not act(b)
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
    // byte offsets of this lane inside a half-row slot: unit (e, chunk cl of the half) adds e * 64 * 16 + cl * 2 * SLAB
    const uint32_t so_hi = sbase + C::OFF_ROWS + g * C::SLAB + (1 + px0) * 16 + q * 8;
    const uint32_t so_p1 = sbase + C::OFF_ROWS + C::HALF_PLANE + (1 + px0) * 16 + co;
    // ring-halo duty (one warp per row, in turns): lane -> side = lane / 16, channel quad cq = lane % 16 (half cq / 8)
    const int h_side = lane >> 4, h_cq = lane & 15;
    const uint32_t h_pos = (h_side ? OPX - 1 : 0) * 16;
    const uint32_t ho_hi = sbase + C::OFF_ROWS + ((h_cq >> 1) & 3) * C::SLAB + h_pos + (h_cq & 1) * 8;
    const uint32_t ho_p1 = sbase + C::OFF_ROWS + C::HALF_PLANE + ((h_cq >> 2) & 1) * 2 * C::SLAB + h_pos + (h_cq & 3) * 4;

    float4 ring[8], hring = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int u = 0; u < 8; ++u) ring[u] = make_float4(0.f, 0.f, 0.f, 0.f);
    const float* l_ptr = nullptr;      // this lane's float4 of unit (e = 0, c = 0) of the row being loaded
    const float* l_hptr = nullptr;
    bool l_on = false, l_h = false;
    auto load_ctx = [&](const ColRows& r, uint32_t ir_of) {
        l_on = r.valid();
        l_ptr = p.x0 + ((size_t)(r.b * p.H + r.rho) * p.W + r.wt * PIX + px0) * C::CIN + co;
        int ww = r.wt * PIX + (h_side ? PIX : -1);
        bool ok = l_on && (int)(ir_of & 7u) == tw;
        if (ww < 0) { ww += p.W; ok = ok && p.ring; }
        else if (ww >= p.W) { ww -= p.W; ok = ok && p.ring; }
        l_h = ok;
        l_hptr = p.x0 + ((size_t)(r.b * p.H + r.rho) * p.W + ww) * C::CIN + h_cq * 4;
    };
    auto load_unit = [&](int u) {      // u = e * 4 + c (compile-time)
        if (l_on && !ABL(64)) ring[u] = ldg_stream_f4(l_ptr + (u >> 2) * 64 * C::CIN + (u & 3) * 16);
    };

    unsigned long long* dbg = (tw == 0 && lane == 0) ? g_conv_dbg : nullptr;
    unsigned long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const long long t_start = dbg ? clock64() : 0;
    ColRows cur;
    cur.init(t_lo, t_hi, p.H, WT);
    uint32_t ir = 0;
    volatile uint32_t* rows_done = reinterpret_cast<volatile uint32_t*>(__cvta_shared_to_generic(bar0 + 232u));
    load_ctx(cur, 0);
#pragma unroll
    for (int u = 0; u < 8; ++u) load_unit(u);
    if (l_h) hring = ldg_stream_f4(l_hptr);
    mbar_wait_quiet(bar0 + 200u, 0);     // COEF_FULL
    while (cur.valid()) {
        ColRows nx = cur;
        nx.next();
        const float* sa = s_coef + (cur.b - b_lo) * 64;
        const float* sb = sa + 128;
        const bool halo_duty = (int)(ir & 7u) == tw;
        // this row's ring-halo quads (both halves) are converted up front; each half is stored with its half slot
        uint32_t hw0 = 0, hw1 = 0, hw2 = 0, hw3 = 0;
        {
            const float4 x4 = hring;
            load_ctx(nx, ir + 1);
            if (l_h) hring = ldg_stream_f4(l_hptr);
            if (halo_duty) {
                const int ww = cur.wt * PIX + (h_side ? PIX : -1);
                const bool valid = p.ring || (ww >= 0 && ww < p.W);
                const float4 a4 = *reinterpret_cast<const float4*>(sa + h_cq * 4);
                const float4 b4 = *reinterpret_cast<const float4*>(sb + h_cq * 4);
                const float ha[4] = {a4.x, a4.y, a4.z, a4.w}, hb[4] = {b4.x, b4.y, b4.z, b4.w};
                col_convert<SILU>(x4, ha, hb, valid, hw0, hw1, hw2, hw3);
            }
        }
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const uint32_t hs = (2u * ir + half) % C::NHALF;
            {
                DBG_T0();
                mbar_wait_quiet(bar0 + 32u + 8u * hs, ((ir >> 1) & 1) ^ 1);      // EMPTY_HALF(hs)
                DBG_ACC(6);
            }
            const uint32_t st = hs * C::HALF;
            if (!ABL(128)) {
#pragma unroll
                for (int cl = 0; cl < 2; ++cl) {
                    const int c = half * 2 + cl;
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
                        const uint32_t off = st + e * 64 * 16 + cl * 2 * C::SLAB;
                        uint32_t w0, w1, w2, w3;
                        col_convert<SILU>(x[e], ca, cb, true, w0, w1, w2, w3);
                        if (!ABL(256)) {
                            sts_v2(so_hi + off, w0, w1);
                            sts_b32(so_p1 + off, w2);
                            sts_b32(so_p1 + off + C::SLAB, w3);
                        }
                    }
                }
                if (halo_duty && (h_cq >> 3) == half) {
                    sts_v2(ho_hi + st, hw0, hw1);
                    sts_b32(ho_p1 + st, hw2);
                    sts_b32(ho_p1 + st + C::SLAB, hw3);
                }
            }
            // The generic-proxy -> async-proxy fence for these stores is executed by the CONSUMER (the MMA issuer warp, after its
            // acquire of FULL_HALF): here ptxas lowers fence.proxy.async to MEMBAR.ALL.CTA + FENCE.VIEW.ASYNC, and the MEMBAR
            // would also wait for this warp's prefetch loads of the NEXT row (one exposed HBM round trip per row).  The arrive
            // below is a release at CTA scope: the stores are performed before the phase completes.
#if B200_COL_FENCE_PRODUCER
            fence_proxy_async();
#endif
            __syncwarp();
            if (lane == 0) mbar_arrive(bar0 + 8u * hs);                           // FULL_HALF(hs)
        }
        ++ir;
        cur = nx;
        if (tw == 0 && lane == 0) *rows_done = ir;
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
#define CFULL_HALF(s) (bar0 + 8u * (s))
#define CEMPTY_HALF(s) (bar0 + 32u + 8u * (s))
#define CACC_FULL(s) (bar0 + 64u + 8u * (s))
#define CACC_EMPTY(s) (bar0 + 128u + 8u * (s))
#define CW_FULL (bar0 + 192u)
#define CCOEF_FULL (bar0 + 200u)

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int WT = p.W / PIX;
    const int per_cta = (p.n_tiles + gridDim.x - 1) / gridDim.x;
    const int t_lo = blockIdx.x * per_cta;
    const int t_hi = min(t_lo + per_cta, p.n_tiles);
    const int per_sample = WT * p.H;
    const int b_lo = t_lo / per_sample, b_hi = (max(t_hi, t_lo + 1) - 1) / per_sample;   // <= b_lo + 1 (checked on the host)

    if (threadIdx.x == 0) {
        for (int s = 0; s < C::NHALF; ++s) {
            mbar_init(CFULL_HALF(s), XF_WARPS);
            mbar_init(CEMPTY_HALF(s), 1);          // read by ONE issuer warp (warp 5: channels 0-31, warp 7: channels 32-63)
        }
        for (int s = 0; s < C::NACC; ++s) {
            mbar_init(CACC_FULL(s), 2);            // one commit from each MMA issuer warp
            mbar_init(CACC_EMPTY(s), C::EW * 32);
        }
        mbar_init(CW_FULL, 1);
        mbar_init(CCOEF_FULL, 1);
        *reinterpret_cast<volatile uint32_t*>(smem + C::OFF_BAR + 232) = 0u;     // input rows converted so far (prefetch pacing)
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
    // accumulator of the CTA's i-th output row: ring of 8 x 64 TMEM columns in DESCENDING order
    auto slot_of = [](uint32_t it) { return (8u - (it & 7u)) & 7u; };
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
            // ------------------------------ weights (once, resident) + L2 prefetch of the rows ahead of the transform warps ------------------------------
            if (lane == 0 && t_lo < t_hi) {
                const uint8_t* wsrc = reinterpret_cast<const uint8_t*>(p.w);
                mbar_expect_tx(CW_FULL, C::WALL);
                for (int i = 0; i < C::WALL / C::WIMG; ++i)
                    bulk_copy_g2s(sbase + C::OFF_W + i * C::WIMG, wsrc + (size_t)i * C::WIMG, C::WIMG, CW_FULL);
                constexpr int PF = 4;                         // prefetch distance in input rows
                constexpr uint32_t ROWB = PIX * C::CIN * 4;
                ColRows pr;
                pr.init(t_lo, t_hi, p.H, WT);
                auto prefetch = [&](const ColRows& r) {
                    const size_t col = ((size_t)r.b * p.H * p.W + r.wt * PIX) * C::CIN;
                    l2_prefetch_bulk(p.x0 + col + (size_t)r.rho * p.W * C::CIN, ROWB);
                    if (p.res && r.rho >= r.h_a && r.rho <= r.h_b) l2_prefetch_bulk(p.res + col + (size_t)r.rho * p.W * C::CIN, ROWB);
                };
                for (int i = 0; i < PF && pr.valid(); ++i, pr.next()) prefetch(pr);
                // paced by the transform warps through a monotonic row counter (a parity wait on FULL_HALF could miss a phase if
                // this thread ever fell two rows behind): when input row ir is complete, request row ir + PF
                const volatile uint32_t* rows_done = reinterpret_cast<const volatile uint32_t*>(smem + C::OFF_BAR + 232);
                for (uint32_t ir = 0; pr.valid(); ++ir, pr.next()) {
                    while (*rows_done <= ir) __nanosleep(256);
                    prefetch(pr);
                }
            }
        } else {
            // ------------------------------ MMA issuers (warps 5 / 7): issue unit = (input row, half of the channels) ------------------------------
            // = 2 chunks x 3 dx x (fp16 + e4m3) = 12 MMAs of N = 192 (~1150 cycles of tensor pipe).  Warp 5 takes channels 0-31,
            // warp 7 channels 32-63 of every row: they alternate on the tensor pipe (named barriers 2 / 3, as in the tile-walk
            // kernel): while one warp's MMAs run, the other has already polled the barriers of its next unit.
            // tcgen05.commit covers the MMAs of the EXECUTING thread only: a half-row slot is read by one warp (1 arrival), an
            // accumulator is written by both (2 arrivals: each warp commits after its unit of the last contributing row).
            if (t_lo < t_hi) {
                const uint32_t par = warp == 7 ? 1u : 0u;
                constexpr uint32_t idesc0 = make_idesc_f16(128, 0);
                unsigned long long* dbg = (lane == 0 && par == 0) ? g_conv_dbg : nullptr;
                unsigned long long dbg_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                const long long t_start = dbg ? clock64() : 0;
                mbar_wait(CW_FULL, 0);
                ColRows r;
                r.init(t_lo, t_hi, p.H, WT);
                uint32_t opened = 0, closed = 0;      // accumulators (output rows of the CTA) opened / handed to the epilogue
                for (uint32_t ir = 0; r.valid(); ++ir, r.next()) {
                    const int o_hi = min(r.rho + 1, r.h_b), o_lo = max(r.rho - 1, r.h_a);
                    const uint32_t n_out = (uint32_t)(o_hi - o_lo + 1);
                    const uint32_t dy_lo = (uint32_t)(r.rho - o_hi + 1);
                    const uint32_t it_hi = (uint32_t)(r.it_a + (o_hi - r.h_a));
                    const uint32_t s0 = slot_of(it_hi);
                    const uint32_t run1 = min(n_out, 8u - s0), run2 = n_out - run1;     // accumulator ring wrap-around
                    if (opened <= it_hi) {
                        DBG_T0();
                        for (; opened <= it_hi; ++opened) mbar_wait(CACC_EMPTY(slot_of(opened)), (opened >> 3) & 1);
                        DBG_ACC(3);
                    }
                    const uint32_t hs = (2u * ir + par) % C::NHALF;
                    {
                        DBG_T0();
                        mbar_wait_sleepy<64>(CFULL_HALF(hs), (ir >> 1) & 1);
                        DBG_ACC(1);
                    }
#if !B200_COL_FENCE_PRODUCER
                    fence_proxy_async();     // the transform warps' st.shared (acquired above) -> async-proxy operand reads
#endif
                    const uint32_t a16 = desc_lo(sbase + C::OFF_ROWS + hs * C::HALF, C::SLAB);
                    const uint32_t bw = desc_lo(sbase + C::OFF_W + (2 * par) * 3 * 2 * C::WIMG + dy_lo * (C::BN * 16), C::WROWS * 16);
                    const uint32_t d1 = tmem_base + s0 * C::BN;
                    const uint32_t id1 = idesc0 + ((8u * run1) << 17), id2 = idesc0 + ((8u * run2) << 17);
                    if (ir != 0 || par != 0) {
                        DBG_T0();
                        named_bar_sync(2 + par, 64);
                        DBG_ACC(4);
                    }
                    tc_fence_after();
                    const int done_hi = r.rho == r.rho_b ? r.h_b : r.rho - 1;      // output rows complete after this input row
                    const uint32_t it_done = (uint32_t)(r.it_a + (done_hi - r.h_a));
                    if (elect_one()) {
                        if (!ABL(8)) {
#pragma unroll
                            for (int cl = 0; cl < 2; ++cl) {
#pragma unroll
                                for (int dx = 0; dx < 3; ++dx) {
                                    const uint32_t ao = (uint32_t)((cl * 2 * C::SLAB) >> 4) + dx;
                                    const uint32_t bo = (uint32_t)(((cl * 3 + dx) * 2 * C::WIMG) >> 4);
                                    tc_mma_f16_lh(d1, a16 + ao, bw + bo, id1, 1u);
                                    tc_mma_f8_lh(d1, a16 + ao + (C::HALF_PLANE >> 4), bw + bo + (C::WIMG >> 4), id1, 1u);
                                    if (run2) {
                                        const uint32_t b2 = bw + bo + ((run1 * C::BN * 16) >> 4);
                                        tc_mma_f16_lh(tmem_base, a16 + ao, b2, id2, 1u);
                                        tc_mma_f8_lh(tmem_base, a16 + ao + (C::HALF_PLANE >> 4), b2 + (C::WIMG >> 4), id2, 1u);
                                    }
                                }
                            }
                        }
                        tc_commit(CEMPTY_HALF(hs));
                        if (done_hi >= r.h_a)
                            for (uint32_t i = closed; i <= it_done; ++i) tc_commit(CACC_FULL(slot_of(i)));
                    }
                    if (done_hi >= r.h_a) closed = it_done + 1u;
                    __syncwarp();
                    tc_fence_before();
                    named_bar_arrive(2 + (par ^ 1u), 64);
                }
                if (dbg) {
                    dbg[blockIdx.x * 8 + 0] = clock64() - t_start;
                    dbg[blockIdx.x * 8 + 1] = dbg_acc[1];
                    dbg[blockIdx.x * 8 + 2] = dbg_acc[4];
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
            const uint32_t tq = tmem_base + slice * 32 + ((uint32_t)(quarter * 32) << 16);
            // every MMA accumulates: the accumulators start zeroed, and are handed back zeroed
#pragma unroll 1
            for (int s = 0; s < C::NACC; ++s) {
                tmem_st_zero_32x32(tq + s * C::BN);
                tc_fence_before();
                mbar_arrive(CACC_EMPTY(s));
            }
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
            ColPos cp = col_pos(t_lo, p.H, WT);
            for (int t = t_lo; t < t_hi; ++t, ++it) {
                if (it != 0 && ++cp.h == p.H) {        // next column
                    cp.h = 0;
                    if (++cp.wt == WT) { cp.wt = 0; ++cp.b; }
                }
                if (p.stats && cp.b != cur_b) {
                    if (cur_b >= 0) flush_stats();
                    cur_b = cp.b;
                }
                const uint32_t slot = slot_of(it);
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
                    mbar_wait_sleepy<400>(CACC_FULL(slot), (it >> 3) & 1);
                    DBG_ACC(5);
                }
                tc_fence_after();
                float v[32];
                tmem_ld_32x32(tq + slot * C::BN, v);
                tmem_st_zero_32x32(tq + slot * C::BN);
                tc_fence_before();
                mbar_arrive(CACC_EMPTY(slot));
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
#undef CFULL_HALF
#undef CEMPTY_HALF
#undef CACC_FULL
#undef CACC_EMPTY
#undef CW_FULL
#undef CCOEF_FULL
}

// weight image of the column walk: per (16-channel chunk c, dx) two planes of [2 k-groups][192 = dy * 64 + cout][16 B]:
//   plane 0: fp16(w s), k-group = 8 channels;   plane 1: e4m3, k-group 0 = w s 2^-11, k-group 1 = w s - fp16(w s), 16 channels each
__global__ void pack_weight_col_kernel(const float* __restrict__ w, uint8_t* __restrict__ out, float wscale) {
    using C = ColCfg;
    const int total = C::BN * C::CIN * 9;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        int r = i;
        const int e16 = r % 16; r /= 16;
        const int co = r % C::BN; r /= C::BN;
        const int tap = r % 9; r /= 9;
        const int c = r;
        const int dy = tap / 3, dx = tap % 3, k = c * 16 + e16;
        const float v = w[((size_t)co * C::CIN + k) * 9 + tap] * wscale;
        const __half hi = __float2half_rn(v);
        uint8_t* img = out + (size_t)((c * 3 + dx) * 2) * C::WIMG;
        const int row = dy * C::BN + co;
        *reinterpret_cast<__half*>(img + (e16 >> 3) * (C::WROWS * 16) + row * 16 + (e16 & 7) * 2) = hi;
        img[C::WIMG + row * 16 + e16] = f8x1(v * (1.f / F8_LO_SCALE));
        img[C::WIMG + C::WROWS * 16 + row * 16 + e16] = f8x1(v - __half2float(hi));
    }
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
