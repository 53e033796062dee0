// HBM-bound kernels of the denoiser step: GroupNorm apply + SiLU + fp16 cast (+concat), channel statistics,
// FIR resampling, time embedding, first/last convs on CUDA cores, sampler update.  All NHWC fp32.
#include <stdarg.h>
#include <stdlib.h>

#include "common.cuh"

namespace b200 {

static thread_local char g_err[512] = "";
// B200_PDL=1: every step kernel is launched as a programmatic dependent; B200_PDL=2: only the conv kernels (their
// prologue -- barrier init, TMEM allocation -- and the weight prefetch do not depend on the previous kernel)
static int pdl_mode() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("B200_PDL");
        // measured on B200 inside the step graph (profiles/r01s2_pdl.txt): all kernels 4.01 ms, none 3.92 ms, conv only
        // 3.84-3.89 ms -> default 2
        v = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 2;
    }
    return v;
}
bool pdl_enabled() { return pdl_mode() == 1; }
bool pdl_enabled_conv() { return pdl_mode() >= 1; }

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// ---------------------------------------------------------------------------------------------------------
// GroupNorm apply (+AdaGN) + SiLU -> fp16, optional concat of two sources
// ---------------------------------------------------------------------------------------------------------
constexpr int GN_MAX_C = 1024;


// Per-channel affine coefficients of GroupNorm(+AdaGN) for sample b:  y = x * s_a[c] + s_b[c].
// All threads load the per-channel {sum, sumsq} in parallel (the serial per-group loop of the first version cost
// ~10 us of dependent DRAM latency per block and dominated the kernel).
__device__ __forceinline__ void gn_coefficients(const double* __restrict__ st0, int C0, const double* __restrict__ st1,
                                                int C1, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                const float* __restrict__ ada, int ada_stride, int groups, float eps,
                                                int HW, int b, float* s_a, float* s_b, double* s_st, float* s_mean,
                                                float* s_rstd) {
    const int C = C0 + C1;
    const int cpg = C / groups;
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const double* p = c < C0 ? st0 + ((size_t)b * C0 + c) * 2 : st1 + ((size_t)b * C1 + (c - C0)) * 2;
        s_st[2 * c] = p[0];
        s_st[2 * c + 1] = p[1];
    }
    __syncthreads();
    if ((int)threadIdx.x < groups) {
        double s = 0.0, ss = 0.0;
        for (int c = threadIdx.x * cpg; c < (threadIdx.x + 1) * cpg; ++c) {
            s += s_st[2 * c];
            ss += s_st[2 * c + 1];
        }
        const double n = (double)HW * cpg;
        const double mean = s / n;
        double var = ss / n - mean * mean;
        if (var < 0.0) var = 0.0;
        s_mean[threadIdx.x] = (float)mean;
        s_rstd[threadIdx.x] = (float)(1.0 / sqrt(var + (double)eps));
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        const int g = c / cpg;
        float a = s_rstd[g], bb = -s_mean[g] * s_rstd[g];
        if (gamma) { a *= gamma[c]; bb = bb * gamma[c] + beta[c]; }
        if (ada) {
            const float sc = 1.f + ada[(size_t)b * ada_stride + c];
            const float sh = ada[(size_t)b * ada_stride + C + c];
            a *= sc;
            bb = bb * sc + sh;
        }
        s_a[c] = a;
        s_b[c] = bb;
    }
}

// output layout: the conv operand of common.cuh ("tile-major slabs"): y[plane][b][h][W/128][C/8][130][8]
//
// Work decomposition: block = (sample b, channel block of CB channels, a strided set of 128-pixel tiles); the grid is one
// resident wave (~4 blocks / SM).  Warp item = 8 consecutive pixels x 32 consecutive channels (lane -> pixel lane/4,
// 8-channel group lane%4): reads move whole 128-byte lines (32 channels of one pixel), writes move 128 contiguous
// bytes of a slab (8 pixels of one channel group).  Four items (8 x 16-byte loads per lane) are requested before any
// math; the first batch is requested before the coefficient prologue so that its latency overlaps it.
struct GnItem {
    const float* src;     // this lane's 8 input channels
    size_t oi, oi2;       // fp16-element offsets of the 16-byte output unit and of its halo duplicate
    int c;                // first channel (relative to the block's channel block: index into s_a / s_b)
    bool ok, dup;
};

__global__ void __launch_bounds__(256, 4) gn_act_kernel(const float* __restrict__ x0, int C0, const float* __restrict__ x1,
                                                     int C1, const double* __restrict__ st0,
                                                     const double* __restrict__ st1, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, const float* __restrict__ ada,
                                                     int ada_stride, int groups, float eps, int silu,
                                                     __half* __restrict__ y, __half* __restrict__ y_raw, size_t lo_off,
                                                     int parts, int H, int W, int CB, int tile_blocks) {
    extern __shared__ __align__(16) unsigned char gn_smem[];
    pdl_launch_dependents();
    pdl_wait();
    const int C = C0 + C1, c8n = C / 8;
    const int HW = H * W, WT = W / OTW, tiles = H * WT;
    const int b = blockIdx.y;
    const int cbi = blockIdx.x / tile_blocks, tb = blockIdx.x - cbi * tile_blocks;
    const int c0 = cbi * CB;
    const int cb_n = min(CB, C - c0);                       // channels of this block
    float* s_a = reinterpret_cast<float*>(gn_smem);         // [CB]
    float* s_b = s_a + CB;                                  // [CB]
    double* s_st = reinterpret_cast<double*>(gn_smem + 8 * CB);   // [2 * span] channel statistics of the covering groups
    __shared__ float s_mean[64], s_rstd[64];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int px = lane >> 2, g = lane & 3;
    const int qn = (cb_n + 31) / 32;                        // 32-channel quads per tile row group
    const int n_it = 16 * qn;                               // warp items per tile

    // item `it` (warp-uniform) of tile t: quad q = it / 16 (32 channels), pixel group pg = it % 16 (8 pixels)
    auto decode = [&](int t, int it) {
        GnItem r;
        r.dup = false;
        r.src = nullptr; r.oi = r.oi2 = 0;
        const int q = it >> 4, pg = it & 15;
        r.c = q * 32 + g * 8;
        r.ok = it < n_it && t < tiles && r.c < cb_n;
        if (!r.ok) return r;
        const int hh = t / WT, wt = t - hh * WT;
        const int c = c0 + r.c, c8 = c >> 3;
        const int tp = pg * 8 + px;                          // pixel inside the tile
        const size_t pix = (size_t)b * HW + (size_t)hh * W + wt * OTW + tp;
        r.src = c < C0 ? x0 + pix * C0 + c : x1 + pix * C1 + (c - C0);
        const size_t bh = (size_t)b * H + hh;
        r.oi = operand_unit(bh, WT, c8n, wt, c8, tp + 1) * 8;
        if (tp == 0) { r.dup = true; r.oi2 = operand_unit(bh, WT, c8n, wt == 0 ? WT - 1 : wt - 1, c8, OPX - 1) * 8; }
        else if (tp == OTW - 1) { r.dup = true; r.oi2 = operand_unit(bh, WT, c8n, wt == WT - 1 ? 0 : wt + 1, c8, 0) * 8; }
        return r;
    };
    constexpr int NU = 4;      // items in flight per warp: 8 x 16-byte loads per lane before any math
    float4 ld[NU][2];
    auto load_batch = [&](int t, int it0) {
#pragma unroll
        for (int u = 0; u < NU; ++u) {
            const GnItem r = decode(t, it0 + u * nwarps);
            if (r.ok) {
                ld[u][0] = *reinterpret_cast<const float4*>(r.src);
                ld[u][1] = *reinterpret_cast<const float4*>(r.src + 4);
            }
        }
    };

    // ---- first loads in flight before the coefficient prologue ----
    load_batch(tb, warp);

    // ---- per-channel affine coefficients of GroupNorm(+AdaGN) for sample b: y = x * s_a[c] + s_b[c] ----
    if (st0 != nullptr) {
        const int cpg = C / groups;
        const int g_lo = c0 / cpg, g_hi = (c0 + cb_n - 1) / cpg;          // groups overlapping this channel block
        const int span0 = g_lo * cpg, span = (g_hi + 1) * cpg - span0;   // their channels
        for (int i = threadIdx.x; i < span; i += blockDim.x) {
            const int c = span0 + i;
            const double* p = c < C0 ? st0 + ((size_t)b * C0 + c) * 2 : st1 + ((size_t)b * C1 + (c - C0)) * 2;
            s_st[2 * i] = p[0];
            s_st[2 * i + 1] = p[1];
        }
        // own coefficients' inputs in the same memory round trip
        float ga = 1.f, be = 0.f, sc = 1.f, sh = 0.f;
        const int cown = c0 + (int)threadIdx.x;
        const bool own = (int)threadIdx.x < cb_n;
        if (own) {
            if (gamma) { ga = gamma[cown]; be = beta[cown]; }
            if (ada) { sc = 1.f + ada[(size_t)b * ada_stride + cown]; sh = ada[(size_t)b * ada_stride + C + cown]; }
        }
        __syncthreads();
        for (int gi = g_lo + warp; gi <= g_hi; gi += nwarps) {           // one warp per group, lanes over its channels
            double s = 0.0, ss = 0.0;
            for (int i = lane; i < cpg; i += 32) {
                s += s_st[2 * ((gi - g_lo) * cpg + i)];
                ss += s_st[2 * ((gi - g_lo) * cpg + i) + 1];
            }
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s += __shfl_xor_sync(0xffffffffu, s, o);
                ss += __shfl_xor_sync(0xffffffffu, ss, o);
            }
            if (lane == 0) {
                const double n = (double)HW * cpg;
                const double mean = s / n;
                double var = ss / n - mean * mean;
                if (var < 0.0) var = 0.0;
                s_mean[gi - g_lo] = (float)mean;
                s_rstd[gi - g_lo] = (float)(1.0 / sqrt(var + (double)eps));
            }
        }
        __syncthreads();
        if (own) {
            const int gi = cown / cpg - g_lo;
            float a = s_rstd[gi], bb = -s_mean[gi] * s_rstd[gi];
            a *= ga; bb = bb * ga + be;
            a *= sc; bb = bb * sc + sh;
            s_a[threadIdx.x] = a;
            s_b[threadIdx.x] = bb;
        }
    } else {
        for (int c = threadIdx.x; c < cb_n; c += blockDim.x) { s_a[c] = 1.f; s_b[c] = 0.f; }
    }
    __syncthreads();

    bool first = true;
    for (int t = tb; t < tiles; t += tile_blocks) {
        for (int it0 = warp; it0 < n_it; it0 += NU * nwarps) {
            if (!first) load_batch(t, it0);
            first = false;
#pragma unroll
            for (int u = 0; u < NU; ++u) {
                const GnItem it = decode(t, it0 + u * nwarps);
                if (!it.ok) continue;     // warp-uniform for C % 32 == 0 (required by parts == 3: shuffles below)
                float v[8] = {ld[u][0].x, ld[u][0].y, ld[u][0].z, ld[u][0].w, ld[u][1].x, ld[u][1].y, ld[u][1].z, ld[u][1].w};
                __half2 h[4];
                if (y_raw) {   // second output: the un-normalised input as a conv operand (1x1 skip conv of the same block)
#pragma unroll
                    for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
                    const uint4 hv = *reinterpret_cast<const uint4*>(h);
                    *reinterpret_cast<uint4*>(y_raw + it.oi) = hv;
                    if (it.dup) *reinterpret_cast<uint4*>(y_raw + it.oi2) = hv;
                    if (parts >= 2) {
                        const uint4 pv = plane1_value(v, h, parts, lane);
                        *reinterpret_cast<uint4*>(y_raw + lo_off + it.oi) = pv;
                        if (it.dup) *reinterpret_cast<uint4*>(y_raw + lo_off + it.oi2) = pv;
                    }
                }
                const float4 a0 = *reinterpret_cast<const float4*>(s_a + it.c), a1 = *reinterpret_cast<const float4*>(s_a + it.c + 4);
                const float4 b0 = *reinterpret_cast<const float4*>(s_b + it.c), b1 = *reinterpret_cast<const float4*>(s_b + it.c + 4);
                const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
                const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    float tt = fmaf(v[e], av[e], bv[e]);
                    if (silu) tt = silu_f(tt);
                    v[e] = tt;
                }
#pragma unroll
                for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
                const uint4 hv = *reinterpret_cast<const uint4*>(h);
                *reinterpret_cast<uint4*>(y + it.oi) = hv;
                if (it.dup) *reinterpret_cast<uint4*>(y + it.oi2) = hv;
                if (parts >= 2) {   // parts 2: lo = fp16(x - fp32(hi)); parts 3: e4m3 pair plane
                    const uint4 pv = plane1_value(v, h, parts, lane);
                    *reinterpret_cast<uint4*>(y + lo_off + it.oi) = pv;
                    if (it.dup) *reinterpret_cast<uint4*>(y + lo_off + it.oi2) = pv;
                }
            }
        }
    }
}

// decomposition for large tensors: one block = 2^k (<= 256, <= W) consecutive pixels of ONE image row x all channels.
// ncu on B200 showed the first version of this kernel ISSUE-bound (smsp issue active 68 %, ~500 instructions per
// 8-element item, most of them 64-bit address arithmetic and an integer division per item), not memory-bound: here
// everything that can be per-block is (row, base pointers), per-item offsets are 32-bit shifts / multiplies, and
// `parts` / the raw second output are template parameters.
template <int PARTS, bool RAW>
__global__ void __launch_bounds__(256) gn_act_kernel_v1(const float* __restrict__ x0, int C0, const float* __restrict__ x1,
                                                        int C1, const double* __restrict__ st0,
                                                        const double* __restrict__ st1, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, const float* __restrict__ ada,
                                                        int ada_stride, int groups, float eps, int silu,
                                                        __half* __restrict__ y, __half* __restrict__ y_raw, size_t lo_off,
                                                        int HW, int W, int ppb_log2) {
    // dynamic shared memory sized by the channel count (24 B / channel): keeps 8 blocks resident per SM
    extern __shared__ __align__(16) unsigned char gn_smem[];
    pdl_launch_dependents();
    pdl_wait();
    const int C = C0 + C1;
    double* s_st = reinterpret_cast<double*>(gn_smem);            // [2C]
    float* s_a = reinterpret_cast<float*>(gn_smem + 16 * C);      // [C]
    float* s_b = s_a + C;                                         // [C]
    __shared__ float s_mean[64], s_rstd[64];
    const int b = blockIdx.y;
    if (st0 != nullptr) {
        gn_coefficients(st0, C0, st1, C1, gamma, beta, ada, ada_stride, groups, eps, HW, b, s_a, s_b, s_st, s_mean, s_rstd);
    } else {
        for (int c = threadIdx.x; c < C; c += blockDim.x) { s_a[c] = 1.f; s_b[c] = 0.f; }
    }
    __syncthreads();
    const int c8n = C / 8, WT = W >> 7, Himg = HW / W;
    const int p0 = blockIdx.x << ppb_log2;                  // the block lies inside one image row (W % ppb == 0)
    const int hh = p0 / W, w_start = p0 - hh * W;
    const int pbn_log2 = ppb_log2 - 3, pb_mask = (1 << pbn_log2) - 1;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int px = lane >> 2, g = lane & 3;
    const int cb_n = (C + 31) / 32;
    const int n_items = cb_n << pbn_log2;
    // per-block 64-bit bases; everything per item below is 32-bit
    const float* x0b = x0 + ((size_t)b * HW + p0) * C0;
    const float* x1b = C1 ? x1 + ((size_t)b * HW + p0) * C1 : nullptr;
    const size_t row_off = ((size_t)b * Himg + hh) * WT * c8n * (OPX * 8);
    __half* yb = y + row_off;
    __half* yrb = RAW ? y_raw + row_off : nullptr;
    // One warp item = 8 consecutive pixels x 32 consecutive channels: lane -> (pixel lane/4, 8-channel group lane%4).
    // Reads: each pixel's 32 channels are one 128-byte line; writes: for each of the 4 channel groups the 8 pixels
    // are 128 contiguous bytes of the tile-major operand.  Two items per iteration: all four 16-byte loads are
    // issued before any math (memory-level parallelism).
    for (int it0 = warp; it0 < n_items; it0 += 2 * nwarps) {
        float4 ld[2][2];
        int c8s[2], pls[2];
        bool ok[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int it = it0 + u * nwarps;
            c8s[u] = (it >> pbn_log2) * 4 + g;
            pls[u] = ((it & pb_mask) << 3) + px;
            ok[u] = it < n_items && c8s[u] < c8n;
            if (ok[u]) {
                const int c = c8s[u] * 8;
                const float* src = c < C0 ? x0b + pls[u] * C0 + c : x1b + pls[u] * C1 + (c - C0);
                ld[u][0] = *reinterpret_cast<const float4*>(src);
                ld[u][1] = *reinterpret_cast<const float4*>(src + 4);
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            if (!ok[u]) continue;     // warp-uniform for C % 32 == 0 (required by parts == 3: shuffles below)
            const int c8 = c8s[u], c = c8 * 8;
            const int ww = w_start + pls[u], wt = ww >> 7, tp = ww & (OTW - 1);
            const int oi = ((wt * c8n + c8) * OPX + tp + 1) * 8;
            int oi2 = -1;                                         // halo duplicate in the neighbouring tile
            if (tp == 0) oi2 = (((wt == 0 ? WT - 1 : wt - 1) * c8n + c8) * OPX + OPX - 1) * 8;
            else if (tp == OTW - 1) oi2 = (((wt == WT - 1 ? 0 : wt + 1) * c8n + c8) * OPX) * 8;
            float v[8] = {ld[u][0].x, ld[u][0].y, ld[u][0].z, ld[u][0].w, ld[u][1].x, ld[u][1].y, ld[u][1].z, ld[u][1].w};
            __half2 h[4];
            if (RAW) {   // second output: the un-normalised input as a conv operand (1x1 skip conv of the same block)
#pragma unroll
                for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
                const uint4 hv = *reinterpret_cast<const uint4*>(h);
                *reinterpret_cast<uint4*>(yrb + oi) = hv;
                if (oi2 >= 0) *reinterpret_cast<uint4*>(yrb + oi2) = hv;
                if (PARTS >= 2) {
                    const uint4 pv = plane1_value(v, h, PARTS, lane);
                    *reinterpret_cast<uint4*>(yrb + lo_off + oi) = pv;
                    if (oi2 >= 0) *reinterpret_cast<uint4*>(yrb + lo_off + oi2) = pv;
                }
            }
            const float4 a0 = *reinterpret_cast<const float4*>(s_a + c), a1 = *reinterpret_cast<const float4*>(s_a + c + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(s_b + c), b1 = *reinterpret_cast<const float4*>(s_b + c + 4);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                float t = fmaf(v[e], av[e], bv[e]);
                if (silu) t = silu_f(t);
                v[e] = t;
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
            const uint4 hv = *reinterpret_cast<const uint4*>(h);
            *reinterpret_cast<uint4*>(yb + oi) = hv;
            if (oi2 >= 0) *reinterpret_cast<uint4*>(yb + oi2) = hv;
            if (PARTS >= 2) {   // parts 2: lo = fp16(x - fp32(hi)); parts 3: e4m3 pair plane
                const uint4 pv = plane1_value(v, h, PARTS, lane);
                *reinterpret_cast<uint4*>(yb + lo_off + oi) = pv;
                if (oi2 >= 0) *reinterpret_cast<uint4*>(yb + lo_off + oi2) = pv;
            }
        }
    }
}

// fp32 NHWC variant: y = act(GN(x)) kept in fp32 (input of the FIR resampler in up/down ResBlocks,
// layout_unet_v1.py:229-235, and of the final out conv).  Same coefficient prologue as gn_act_kernel.
__global__ void __launch_bounds__(256) gn_act_f32_kernel(const float* __restrict__ x, int C,
                                                         const double* __restrict__ st, const float* __restrict__ gamma,
                                                         const float* __restrict__ beta, int groups, float eps, int silu,
                                                         float* __restrict__ y, double* __restrict__ stats_out, int HW,
                                                         int pix_per_block) {
    extern __shared__ __align__(16) unsigned char gn_smem[];
    pdl_launch_dependents();
    pdl_wait();
    double* s_st = reinterpret_cast<double*>(gn_smem);
    float* s_a = reinterpret_cast<float*>(gn_smem + 16 * C);
    float* s_b = s_a + C;
    __shared__ float s_mean[64], s_rstd[64];
    const int b = blockIdx.y;
    gn_coefficients(st, C, nullptr, 0, gamma, beta, nullptr, 0, groups, eps, HW, b, s_a, s_b, s_st, s_mean, s_rstd);
    __syncthreads();
    const int c4n = C / 4;
    const int p0 = blockIdx.x * pix_per_block;
    const int np = min(pix_per_block, HW - p0);
    for (int i = threadIdx.x; i < np * c4n; i += blockDim.x) {
        const int pp = i / c4n, c = (i - pp * c4n) * 4;
        const size_t gi = ((size_t)b * HW + p0 + pp) * C + c;
        float4 v = *reinterpret_cast<const float4*>(x + gi);
        v.x = fmaf(v.x, s_a[c], s_b[c]); v.y = fmaf(v.y, s_a[c + 1], s_b[c + 1]);
        v.z = fmaf(v.z, s_a[c + 2], s_b[c + 2]); v.w = fmaf(v.w, s_a[c + 3], s_b[c + 3]);
        if (silu) { v.x = silu_f(v.x); v.y = silu_f(v.y); v.z = silu_f(v.z); v.w = silu_f(v.w); }
        *reinterpret_cast<float4*>(y + gi) = v;
    }
}

// ---------------------------------------------------------------------------------------------------------
// per-(b, c) sum / sum of squares of an NHWC fp32 tensor (C/4 must divide 256 or be a multiple of it)
// ---------------------------------------------------------------------------------------------------------
constexpr int ST_PIX_PER_BLOCK = 256;

__device__ __forceinline__ void block_channel_reduce(float4 s1, float4 s2, int C, int b, double* stats,
                                                     float* red /* [256*8] */) {
    // threads with equal (threadIdx.x % (C/4)) hold partial sums of the same 4 channels
    const int c4n = C / 4;
    float* r = red + threadIdx.x * 8;
    r[0] = s1.x; r[1] = s1.y; r[2] = s1.z; r[3] = s1.w;
    r[4] = s2.x; r[5] = s2.y; r[6] = s2.z; r[7] = s2.w;
    __syncthreads();
    if ((int)threadIdx.x < c4n) {
        float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int t = threadIdx.x; t < (int)blockDim.x; t += c4n)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] += red[t * 8 + e];
        double* st = stats + ((size_t)b * C + threadIdx.x * 4) * 2;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            atomicAdd(st + 2 * e, (double)acc[e]);
            atomicAdd(st + 2 * e + 1, (double)acc[4 + e]);
        }
    }
}

__global__ void __launch_bounds__(256) channel_stats_kernel(const float* __restrict__ x, double* __restrict__ stats,
                                                            int HW, int C) {
    __shared__ float red[256 * 8];
    const int b = blockIdx.y;
    const int c4n = C / 4;  // <= 256, divides 256
    const int c4 = threadIdx.x % c4n, poff = threadIdx.x / c4n, pstep = 256 / c4n;
    const int p0 = blockIdx.x * ST_PIX_PER_BLOCK;
    const int p1 = min(p0 + ST_PIX_PER_BLOCK, HW);
    float4 s1 = make_float4(0, 0, 0, 0), s2 = make_float4(0, 0, 0, 0);
    for (int pp = p0 + poff; pp < p1; pp += pstep) {
        const float4 v = *reinterpret_cast<const float4*>(x + ((size_t)b * HW + pp) * C + c4 * 4);
        s1.x += v.x; s1.y += v.y; s1.z += v.z; s1.w += v.w;
        s2.x += v.x * v.x; s2.y += v.y * v.y; s2.z += v.z * v.z; s2.w += v.w * v.w;
    }
    block_channel_reduce(s1, s2, C, b, stats, red);
}

// split-K reduce: the conv epilogue on the partial sums of the K slices (b200_conv_tc_splitk); slices are added in index order.
// The tensors are small (a deep level at small batch: 512 pixels x 512 channels): the grid is sized for ~64 blocks per launch
// (enough loads in flight, few same-address fp64 atomics) and a thread requests 4 pixels x splits 16-byte loads at a time.
__global__ void __launch_bounds__(256) conv_splitk_reduce_kernel(const float* __restrict__ part, int splits, size_t stride,
                                                                 const float* __restrict__ bias, const float* __restrict__ res,
                                                                 float w_inv, float scale, float* __restrict__ out,
                                                                 double* __restrict__ stats, int HW, int C, int pix_per_block) {
    __shared__ float red[256 * 8];
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y;
    const int c4n = C / 4;  // <= 256, divides 256
    const int c4 = threadIdx.x % c4n, poff = threadIdx.x / c4n, pstep = 256 / c4n;
    const int p0 = blockIdx.x * pix_per_block;
    const int p1 = min(p0 + pix_per_block, HW);
    const float4 bi = bias ? *reinterpret_cast<const float4*>(bias + c4 * 4) : make_float4(0, 0, 0, 0);
    float4 s1 = make_float4(0, 0, 0, 0), s2 = make_float4(0, 0, 0, 0);
    constexpr int PB = 4;
    for (int pb = p0 + poff; pb < p1; pb += PB * pstep) {
        float4 v[PB], r[PB];
#pragma unroll
        for (int j = 0; j < PB; ++j) {
            const int pp = pb + j * pstep;
            const size_t gi = ((size_t)b * HW + min(pp, p1 - 1)) * C + c4 * 4;
            v[j] = *reinterpret_cast<const float4*>(part + gi);
            r[j] = res ? *reinterpret_cast<const float4*>(res + gi) : make_float4(0, 0, 0, 0);
        }
        for (int s = 1; s < splits; ++s) {
#pragma unroll
            for (int j = 0; j < PB; ++j) {
                const int pp = pb + j * pstep;
                const size_t gi = ((size_t)b * HW + min(pp, p1 - 1)) * C + c4 * 4;
                const float4 t = *reinterpret_cast<const float4*>(part + (size_t)s * stride + gi);
                v[j].x += t.x; v[j].y += t.y; v[j].z += t.z; v[j].w += t.w;
            }
        }
#pragma unroll
        for (int j = 0; j < PB; ++j) {
            const int pp = pb + j * pstep;
            if (pp >= p1) continue;
            const size_t gi = ((size_t)b * HW + pp) * C + c4 * 4;
            float4 o;
            o.x = (fmaf(v[j].x, w_inv, bi.x) + r[j].x) * scale; o.y = (fmaf(v[j].y, w_inv, bi.y) + r[j].y) * scale;
            o.z = (fmaf(v[j].z, w_inv, bi.z) + r[j].z) * scale; o.w = (fmaf(v[j].w, w_inv, bi.w) + r[j].w) * scale;
            *reinterpret_cast<float4*>(out + gi) = o;
            s1.x += o.x; s1.y += o.y; s1.z += o.z; s1.w += o.w;
            s2.x += o.x * o.x; s2.y += o.y * o.y; s2.z += o.z * o.z; s2.w += o.w * o.w;
        }
    }
    if (stats) block_channel_reduce(s1, s2, C, b, stats, red);
}

int launch_splitk_reduce(const float* part, int splits, size_t stride, const float* bias, const float* res, float w_inv,
                         float scale, float* out, double* stats, int B, int HW, int Cout, void* stream) {
    const int pstep = 256 / (Cout / 4);
    int ppb = (int)(((long long)HW * B + 63) / 64);           // ~64 blocks per launch
    ppb = ((ppb + pstep - 1) / pstep) * pstep;
    if (ppb < pstep) ppb = pstep;
    if (ppb > ST_PIX_PER_BLOCK) ppb = ST_PIX_PER_BLOCK;
    dim3 grid(cdiv(HW, ppb), B);
    launch_pdl(conv_splitk_reduce_kernel, grid, dim3(256), 0, (cudaStream_t)stream, part, splits, stride, bias, res, w_inv, scale,
               out, stats, HW, Cout, ppb);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

// ---------------------------------------------------------------------------------------------------------
// FIR resample ([1,3,3,1] window), circular W / zero H
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void fma4(float4& a, float k, const float4& v) {
    a.x = fmaf(k, v.x, a.x); a.y = fmaf(k, v.y, a.y); a.z = fmaf(k, v.z, a.z); a.w = fmaf(k, v.w, a.w);
}

template <bool UP>
__global__ void __launch_bounds__(256) fir_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                  double* __restrict__ stats, int H, int W, int C, int ring,
                                                  int pix_per_block) {
    __shared__ float red[256 * 8];
    pdl_launch_dependents();
    pdl_wait();
    const int Ho = UP ? 2 * H : H / 2, Wo = UP ? 2 * W : W / 2;
    const int b = blockIdx.y;
    const int c4n = C / 4;
    const int c4 = threadIdx.x % c4n, poff = threadIdx.x / c4n, pstep = 256 / c4n;
    const int p0 = blockIdx.x * pix_per_block;
    const int p1 = min(p0 + pix_per_block, Ho * Wo);
    const float* xb = x + (size_t)b * H * W * C + c4 * 4;
    float4 s1 = make_float4(0, 0, 0, 0), s2 = make_float4(0, 0, 0, 0);
    for (int pp = p0 + poff; pp < p1; pp += pstep) {
        const int oh = pp / Wo, ow = pp - oh * Wo;
        float4 acc = make_float4(0, 0, 0, 0);
        if (UP) {
            // out[2i] = (x[i-1] + 3x[i]) / 4 ; out[2i+1] = (3x[i] + x[i+1]) / 4    (per axis)
            const int ih = oh >> 1, iw = ow >> 1;
            const int hn = (oh & 1) ? ih + 1 : ih - 1;
            int wn = (ow & 1) ? iw + 1 : iw - 1;
            bool wn_ok = true;
            if (wn < 0) { if (ring) wn += W; else wn_ok = false; }
            else if (wn >= W) { if (ring) wn -= W; else wn_ok = false; }
            const bool hn_ok = hn >= 0 && hn < H;
            fma4(acc, 9.f / 16.f, ld4(xb + ((size_t)ih * W + iw) * C));
            if (wn_ok) fma4(acc, 3.f / 16.f, ld4(xb + ((size_t)ih * W + wn) * C));
            if (hn_ok) {
                fma4(acc, 3.f / 16.f, ld4(xb + ((size_t)hn * W + iw) * C));
                if (wn_ok) fma4(acc, 1.f / 16.f, ld4(xb + ((size_t)hn * W + wn) * C));
            }
        } else {
            // out[i] = (x[2i-1] + 3x[2i] + 3x[2i+1] + x[2i+2]) / 8    (per axis)
            const float kk[4] = {0.125f, 0.375f, 0.375f, 0.125f};
#pragma unroll
            for (int a = 0; a < 4; ++a) {
                const int ih = 2 * oh - 1 + a;
                if (ih < 0 || ih >= H) continue;
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    int iw = 2 * ow - 1 + c;
                    if (iw < 0) { if (ring) iw += W; else continue; }
                    else if (iw >= W) { if (ring) iw -= W; else continue; }
                    fma4(acc, kk[a] * kk[c], ld4(xb + ((size_t)ih * W + iw) * C));
                }
            }
        }
        *reinterpret_cast<float4*>(y + ((size_t)b * Ho * Wo + pp) * C + c4 * 4) = acc;
        s1.x += acc.x; s1.y += acc.y; s1.z += acc.z; s1.w += acc.w;
        s2.x += acc.x * acc.x; s2.y += acc.y * acc.y; s2.z += acc.z * acc.z; s2.w += acc.w * acc.w;
    }
    if (stats) block_channel_reduce(s1, s2, C, b, stats, red);
}

// FIR x1/2 downsample, two horizontally adjacent outputs per thread: their 4 x 4 input windows share two columns, so the
// pair needs 24 instead of 32 16-byte loads (the one-output version was load-issue bound: 16 loads per 16 fma4).  Per output
// the taps are accumulated in the same order with the same weights as fir_kernel<false> (bit-identical results).
__global__ void __launch_bounds__(256) fir_down_pair_kernel(const float* __restrict__ x, float* __restrict__ y,
                                                            double* __restrict__ stats, int H, int W, int C, int ring,
                                                            int pairs_per_block) {
    __shared__ float red[256 * 8];
    pdl_launch_dependents();
    pdl_wait();
    const int Ho = H / 2, Wo = W / 2, Wp = Wo / 2;           // Wo is even (host check)
    const int b = blockIdx.y;
    const int c4n = C / 4;
    const int c4 = threadIdx.x % c4n, poff = threadIdx.x / c4n, pstep = 256 / c4n;
    const int p0 = blockIdx.x * pairs_per_block;
    const int p1 = min(p0 + pairs_per_block, Ho * Wp);
    const float* xb = x + (size_t)b * H * W * C + c4 * 4;
    float4 s1 = make_float4(0, 0, 0, 0), s2 = make_float4(0, 0, 0, 0);
    const float kk[4] = {0.125f, 0.375f, 0.375f, 0.125f};
    for (int pp = p0 + poff; pp < p1; pp += pstep) {
        const int oh = pp / Wp, ow = (pp - oh * Wp) * 2;
        float4 acc0 = make_float4(0, 0, 0, 0), acc1 = make_float4(0, 0, 0, 0);
        int iws[6];
        bool ok[6];
#pragma unroll
        for (int c = 0; c < 6; ++c) {
            int iw = 2 * ow - 1 + c;
            ok[c] = true;
            if (iw < 0) { if (ring) iw += W; else ok[c] = false; }
            else if (iw >= W) { if (ring) iw -= W; else ok[c] = false; }
            iws[c] = iw;
        }
#pragma unroll
        for (int a = 0; a < 4; ++a) {
            const int ih = 2 * oh - 1 + a;
            if (ih < 0 || ih >= H) continue;
            float4 v[6];
#pragma unroll
            for (int c = 0; c < 6; ++c) v[c] = ok[c] ? ld4(xb + ((size_t)ih * W + iws[c]) * C) : make_float4(0, 0, 0, 0);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                if (ok[c]) fma4(acc0, kk[a] * kk[c], v[c]);
                if (ok[c + 2]) fma4(acc1, kk[a] * kk[c], v[c + 2]);
            }
        }
        float* yo = y + ((size_t)b * Ho * Wo + (size_t)oh * Wo + ow) * C + c4 * 4;
        *reinterpret_cast<float4*>(yo) = acc0;
        *reinterpret_cast<float4*>(yo + C) = acc1;
        s1.x += acc0.x + acc1.x; s1.y += acc0.y + acc1.y; s1.z += acc0.z + acc1.z; s1.w += acc0.w + acc1.w;
        s2.x += acc0.x * acc0.x + acc1.x * acc1.x; s2.y += acc0.y * acc0.y + acc1.y * acc1.y;
        s2.z += acc0.z * acc0.z + acc1.z * acc1.z; s2.w += acc0.w * acc0.w + acc1.w * acc1.w;
    }
    if (stats) block_channel_reduce(s1, s2, C, b, stats, red);
}

// FIR x2 upsample written DIRECTLY as the next conv's operand (Block.forward: Resample(up) -> ring conv,
// efficient_unet.py:176-190): same taps / summation order as fir_kernel<true>, no fp32 round trip through HBM and no
// separate cast launch.  Block = one 128-pixel tile of one output row; warp item = 8 pixels x 32 channels.
__global__ void __launch_bounds__(256) fir_up_operand_kernel(const float* __restrict__ x, __half* __restrict__ y,
                                                             size_t lo_off, int parts, int H, int W, int C, int ring) {
    pdl_launch_dependents();
    pdl_wait();
    const int Ho = 2 * H, Wo = 2 * W, WT = Wo / OTW, c8n = C / 8;
    const int b = blockIdx.y;
    const int oh = blockIdx.x / WT, wt = blockIdx.x - oh * WT;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
    const int px = lane >> 2, g = lane & 3;
    const int ih = oh >> 1;
    const int hn = (oh & 1) ? ih + 1 : ih - 1;
    const bool hn_ok = hn >= 0 && hn < H;
    const int qn = (C + 31) / 32;
    const float* xb = x + (size_t)b * H * W * C;
    const size_t bh = (size_t)b * Ho + oh;
    for (int it = warp; it < 16 * qn; it += nwarps) {
        const int q = it >> 4, pg = it & 15;
        const int c = q * 32 + g * 8;
        if (c >= C) continue;                 // warp-uniform when C % 32 == 0 (required for parts == 3)
        const int tp = pg * 8 + px, ow = wt * OTW + tp;
        const int iw = ow >> 1;
        int wn = (ow & 1) ? iw + 1 : iw - 1;
        bool wn_ok = true;
        if (wn < 0) { if (ring) wn += W; else wn_ok = false; }
        else if (wn >= W) { if (ring) wn -= W; else wn_ok = false; }
        float v[8];
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {
            const float* xc = xb + c + 4 * hf;
            float4 acc = make_float4(0, 0, 0, 0);
            fma4(acc, 9.f / 16.f, ld4(xc + ((size_t)ih * W + iw) * C));
            if (wn_ok) fma4(acc, 3.f / 16.f, ld4(xc + ((size_t)ih * W + wn) * C));
            if (hn_ok) {
                fma4(acc, 3.f / 16.f, ld4(xc + ((size_t)hn * W + iw) * C));
                if (wn_ok) fma4(acc, 1.f / 16.f, ld4(xc + ((size_t)hn * W + wn) * C));
            }
            v[4 * hf] = acc.x; v[4 * hf + 1] = acc.y; v[4 * hf + 2] = acc.z; v[4 * hf + 3] = acc.w;
        }
        const int c8 = c >> 3;
        const size_t oi = operand_unit(bh, WT, c8n, wt, c8, tp + 1) * 8;
        size_t oi2 = 0;
        bool dup = false;
        if (tp == 0) { dup = true; oi2 = operand_unit(bh, WT, c8n, wt == 0 ? WT - 1 : wt - 1, c8, OPX - 1) * 8; }
        else if (tp == OTW - 1) { dup = true; oi2 = operand_unit(bh, WT, c8n, wt == WT - 1 ? 0 : wt + 1, c8, 0) * 8; }
        const float one[8] = {1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f, 1.f}, zero[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        gn_apply_store(v, one, zero, 0, parts, y, lo_off, oi, oi2, dup, lane);
    }
}

// ---------------------------------------------------------------------------------------------------------
// time embedding MLP + stacked (scale, shift) projections
// ---------------------------------------------------------------------------------------------------------
// grid (B, E / TEMB_EO): every block recomputes the small first layer (all loads of a thread are independent and
// issued together) and produces TEMB_EO outputs of the second layer, one warp per 4 outputs with all row loads in
// flight before the reductions.  (The first version -- one block per sample, outputs looped serially -- took 38 us
// of pure dependent L2 latency on B200.)
constexpr int TEMB_EO = 32;
__global__ void __launch_bounds__(256) temb_kernel(const float* __restrict__ t, const float* __restrict__ w1,
                                                   const float* __restrict__ b1, const float* __restrict__ w2,
                                                   const float* __restrict__ b2, const float* __restrict__ add,
                                                   float* __restrict__ temb, int Cs, int E) {
    extern __shared__ float sm[];
    pdl_launch_dependents();
    pdl_wait();
    float* e0 = sm;        // [Cs]
    float* h1 = sm + Cs;   // [E]
    const int b = blockIdx.x, o0 = blockIdx.y * TEMB_EO;
    const float tv = t[b];
    const int half = Cs / 2;
    for (int i = threadIdx.x; i < half; i += blockDim.x) {
        const float f = expf(-logf(10000.f) / (float)(half - 1) * (float)i);
        const float a = tv * f;
        e0[i] = sinf(a);
        e0[half + i] = cosf(a);
    }
    __syncthreads();
    for (int o = threadIdx.x; o < E; o += blockDim.x) {
        float acc = b1[o];
        const float* wr = w1 + (size_t)o * Cs;
        if (Cs % 16 == 0) {
            for (int k0 = 0; k0 < Cs; k0 += 16) {
                float4 wv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) wv[u] = ld4(wr + k0 + 4 * u);
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    acc = fmaf(wv[u].x, e0[k0 + 4 * u], acc); acc = fmaf(wv[u].y, e0[k0 + 4 * u + 1], acc);
                    acc = fmaf(wv[u].z, e0[k0 + 4 * u + 2], acc); acc = fmaf(wv[u].w, e0[k0 + 4 * u + 3], acc);
                }
            }
        } else {
            for (int k = 0; k < Cs; ++k) acc = fmaf(wr[k], e0[k], acc);
        }
        h1[o] = silu_f(acc);
    }
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int OPW = TEMB_EO / 8;                    // outputs per warp
    float acc[OPW];
#pragma unroll
    for (int u = 0; u < OPW; ++u) {
        const int o = o0 + warp * OPW + u;
        acc[u] = 0.f;
        if (o < E) {
            const float* wr = w2 + (size_t)o * E;
            for (int k = lane; k < E; k += 32) acc[u] = fmaf(wr[k], h1[k], acc[u]);
        }
    }
#pragma unroll
    for (int u = 0; u < OPW; ++u) {
        const int o = o0 + warp * OPW + u;
        const float v = warp_sum(acc[u]);
        if (lane == 0 && o < E) temb[(size_t)b * E + o] = v + b2[o] + (add ? add[(size_t)b * E + o] : 0.f);
    }
}

constexpr int ADA_ROWS_PER_BLOCK = 8;    // one projection row per warp: 960 blocks for P = 7680, one memory round trip
constexpr int ADA_MAX_B = 16;

__global__ void __launch_bounds__(256) ada_proj_kernel(const float* __restrict__ temb, const float* __restrict__ wp,
                                                       const float* __restrict__ bp, float* __restrict__ ada, int B,
                                                       int E, int P) {
    extern __shared__ float sm[];  // silu(temb) [B][E]
    pdl_launch_dependents();
    pdl_wait();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * ADA_ROWS_PER_BLOCK + warp;
    // this warp's weight row is requested before the activations are staged (independent of them)
    float wv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) wv[u] = (row < P && lane + 32 * u < E) ? wp[(size_t)row * E + lane + 32 * u] : 0.f;
    for (int i = threadIdx.x; i < B * E; i += blockDim.x) sm[i] = silu_f(temb[i]);
    __syncthreads();
    if (row >= P) return;
    float acc[ADA_MAX_B];
#pragma unroll
    for (int b = 0; b < ADA_MAX_B; ++b) acc[b] = 0.f;
    if (E <= 256) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int k = lane + 32 * u;
            if (k < E) {
#pragma unroll
                for (int b = 0; b < ADA_MAX_B; ++b)
                    if (b < B) acc[b] = fmaf(wv[u], sm[b * E + k], acc[b]);
            }
        }
    } else {
        const float* wr = wp + (size_t)row * E;
        for (int k = lane; k < E; k += 32) {
            const float w = wr[k];
#pragma unroll
            for (int b = 0; b < ADA_MAX_B; ++b)
                if (b < B) acc[b] = fmaf(w, sm[b * E + k], acc[b]);
        }
    }
#pragma unroll
    for (int b = 0; b < ADA_MAX_B; ++b) {
        if (b < B) {
            const float v = warp_sum(acc[b]);
            if (lane == 0) ada[(size_t)b * P + row] = v + bp[row];
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// in_conv: few dynamic NCHW channels + precomputed constant part -> NHWC
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) in_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ cst, int cst_batched,
                                                      float* __restrict__ out, double* __restrict__ stats, int H, int W,
                                                      int Cx, int Cout, int ring) {
    __shared__ float red[256 * 8];
    __shared__ float sw[4 * 9 * 256];  // [ci][tap][co] , Cout <= 256
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y;
    for (int i = threadIdx.x; i < Cx * 9 * Cout; i += blockDim.x) {
        const int co = i % Cout, r = i / Cout;
        const int tap = r % 9, ci = r / 9;
        sw[i] = w[((size_t)co * Cx + ci) * 9 + tap];
    }
    __syncthreads();
    const int HW = H * W;
    const int c4n = Cout / 4;
    const int c4 = threadIdx.x % c4n, poff = threadIdx.x / c4n, pstep = 256 / c4n;
    const int p0 = blockIdx.x * ST_PIX_PER_BLOCK;
    const int p1 = min(p0 + ST_PIX_PER_BLOCK, HW);
    float4 s1 = make_float4(0, 0, 0, 0), s2 = make_float4(0, 0, 0, 0);
    for (int pp = p0 + poff; pp < p1; pp += pstep) {
        const int h = pp / W, ww = pp - h * W;
        float4 acc = ld4(cst + ((size_t)(cst_batched ? b : 0) * HW + pp) * Cout + c4 * 4);
        for (int ci = 0; ci < Cx; ++ci) {
            const float* xp = x + ((size_t)b * Cx + ci) * HW;
#pragma unroll
            for (int tap = 0; tap < 9; ++tap) {
                const int gh = h + tap / 3 - 1;
                int gw = ww + tap % 3 - 1;
                if (gh < 0 || gh >= H) continue;
                if (gw < 0) { if (ring) gw += W; else continue; }
                else if (gw >= W) { if (ring) gw -= W; else continue; }
                const float xv = xp[gh * W + gw];
                fma4(acc, xv, *reinterpret_cast<const float4*>(sw + (ci * 9 + tap) * Cout + c4 * 4));
            }
        }
        *reinterpret_cast<float4*>(out + ((size_t)b * HW + pp) * Cout + c4 * 4) = acc;
        s1.x += acc.x; s1.y += acc.y; s1.z += acc.z; s1.w += acc.w;
        s2.x += acc.x * acc.x; s2.y += acc.y * acc.y; s2.z += acc.z * acc.z; s2.w += acc.w * acc.w;
    }
    if (stats) block_channel_reduce(s1, s2, Cout, b, stats, red);
}

// row-tiled variant (W % 128 == 0): one block = 128 pixels of one image row; the 3 input rows (+ halo pixels) of the few
// dynamic channels are staged in shared memory once, so the 18 taps per output are broadcast LDS instead of dependent,
// bounds-checked global loads (4x faster on B200: the kernel is bound by the 64-channel fp32 store, as it should be).
constexpr int IC_PIX = 128;
__global__ void __launch_bounds__(256) in_conv_rows_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                           const float* __restrict__ cst, int cst_batched,
                                                           float* __restrict__ out, double* __restrict__ stats, int H,
                                                           int W, int Cx, int Cout, int ring, int rows_per_block) {
    __shared__ float red[256 * 8];
    __shared__ float sw[4 * 9 * 128];                 // [ci][tap][co], Cout <= 128
    constexpr int IC_RMAX = 4;                        // rows per block (host: rows_per_block <= IC_RMAX)
    __shared__ float sx[4 * (IC_RMAX + 2) * (IC_PIX + 2)];   // [ci][row + 1][pixel + 1]: ALL input rows of the block, staged once
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.z, w0 = blockIdx.x * IC_PIX;
    for (int i = threadIdx.x; i < Cx * 9 * Cout; i += blockDim.x) {
        const int co = i % Cout, r = i / Cout;
        const int tap = r % 9, ci = r / 9;
        sw[i] = w[((size_t)co * Cx + ci) * 9 + tap];
    }
    const int HW = H * W;
    // thread = 4 output channels x (128 / pstep) pixels px = poff + i * pstep: every weight float4 read from shared
    // memory feeds all of the thread's pixels (one-pixel-at-a-time was LDS-bound: ncu l1tex 69 %, 75 us on B200)
    const int c4n = Cout / 4;
    const int c4 = threadIdx.x % c4n, poff = threadIdx.x / c4n, pstep = 256 / c4n;
    float4 s1 = make_float4(0, 0, 0, 0), s2 = make_float4(0, 0, 0, 0);
    constexpr int PB = 8;                              // pixels per register batch
    // several image rows per block: the weight staging above and the statistics reduction below are paid once
    const int h_beg = blockIdx.y * rows_per_block, h_end = min(h_beg + rows_per_block, H);
    // every input row of the block (rows_per_block + 2) in ONE staging pass, issued together with the weight staging above
    // (a pass per output row exposed a global-load round trip + two block barriers per row: 51 -> 47 us at the headline shape)
    const int NRX = rows_per_block + 2;
    for (int i = threadIdx.x; i < Cx * NRX * (IC_PIX + 2); i += blockDim.x) {
        const int px = i % (IC_PIX + 2), r = i / (IC_PIX + 2);
        const int ry = r % NRX, ci = r / NRX;
        const int gh = h_beg + ry - 1;
        int gw = w0 + px - 1;
        bool ok = gh >= 0 && gh < H;
        if (gw < 0) { if (ring) gw += W; else ok = false; }
        else if (gw >= W) { if (ring) gw -= W; else ok = false; }
        sx[i] = ok ? x[((size_t)b * Cx + ci) * HW + (size_t)gh * W + gw] : 0.f;
    }
    __syncthreads();
    for (int h = h_beg; h < h_end; ++h) {
        for (int pb = poff; pb < IC_PIX; pb += PB * pstep) {
            float4 acc[PB];
#pragma unroll
            for (int i = 0; i < PB; ++i) {
                const int px = pb + i * pstep;
                acc[i] = px < IC_PIX ? ld4(cst + ((size_t)(cst_batched ? b : 0) * HW + (size_t)h * W + w0 + px) * Cout + c4 * 4)
                                     : make_float4(0, 0, 0, 0);
            }
            for (int ci = 0; ci < Cx; ++ci) {
#pragma unroll
                for (int tap = 0; tap < 9; ++tap) {
                    const float4 wv = *reinterpret_cast<const float4*>(sw + (ci * 9 + tap) * Cout + c4 * 4);
                    const float* xr = sx + (ci * NRX + (h - h_beg) + tap / 3) * (IC_PIX + 2) + tap % 3;
#pragma unroll
                    for (int i = 0; i < PB; ++i) {
                        const int px = pb + i * pstep;
                        if (px < IC_PIX) fma4(acc[i], xr[px], wv);
                    }
                }
            }
#pragma unroll
            for (int i = 0; i < PB; ++i) {
                const int px = pb + i * pstep;
                if (px >= IC_PIX) continue;
                *reinterpret_cast<float4*>(out + ((size_t)b * HW + (size_t)h * W + w0 + px) * Cout + c4 * 4) = acc[i];
                s1.x += acc[i].x; s1.y += acc[i].y; s1.z += acc[i].z; s1.w += acc[i].w;
                s2.x += acc[i].x * acc[i].x; s2.y += acc[i].y * acc[i].y; s2.z += acc[i].z * acc[i].z; s2.w += acc[i].w * acc[i].w;
            }
        }
    }
    if (stats) block_channel_reduce(s1, s2, Cout, b, stats, red);
}

// generic fp32 direct conv (constant folding only; not a hot kernel)
__global__ void conv_direct_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                   const float* __restrict__ bias, float* __restrict__ out, int B, int H, int W,
                                   int Cin, int Cout, int k, int ring) {
    const long long total = (long long)B * H * W * Cout;
    const int pad = k / 2;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int co = (int)(i % Cout);
        const long long pix = i / Cout;
        const int ww = (int)(pix % W), h = (int)((pix / W) % H), b = (int)(pix / ((long long)W * H));
        float acc = bias ? bias[co] : 0.f;
        for (int dy = 0; dy < k; ++dy) {
            const int gh = h + dy - pad;
            if (gh < 0 || gh >= H) continue;
            for (int dx = 0; dx < k; ++dx) {
                int gw = ww + dx - pad;
                if (gw < 0) { if (ring) gw += W; else continue; }
                else if (gw >= W) { if (ring) gw -= W; else continue; }
                const float* xp = x + ((size_t)(b * H + gh) * W + gw) * Cin;
                const float* wp = w + (size_t)co * Cin * k * k + dy * k + dx;
                for (int ci = 0; ci < Cin; ++ci) acc = fmaf(xp[ci], wp[(size_t)ci * k * k], acc);
            }
        }
        out[i] = acc;
    }
}

// out_conv: NHWC (fp32 or fp16) -> NCHW, Cout <= 4; one thread per pixel, weights broadcast from smem
__device__ __forceinline__ void ld8(const float* p, float* v) {
    const float4 a = ld4(p), b = ld4(p + 4);
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
}
__device__ __forceinline__ void ld8(const __half* p, float* v) {
    const uint4 raw = *reinterpret_cast<const uint4*>(p);
    const __half2* h2 = reinterpret_cast<const __half2*>(&raw);
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 f = __half22float2(h2[e]);
        v[2 * e] = f.x; v[2 * e + 1] = f.y;
    }
}

template <typename TIn>
__global__ void __launch_bounds__(128) out_conv_kernel(const TIn* __restrict__ a, const float* __restrict__ w,
                                                       const float* __restrict__ bias, float* __restrict__ pred, int H,
                                                       int W, int Cin, int Cout, int ring) {
    extern __shared__ float4 sw4[];  // [tap][ci] -> 4 output channels (zero padded)
    float* sw = reinterpret_cast<float*>(sw4);
    for (int i = threadIdx.x; i < 9 * Cin * 4; i += blockDim.x) {
        const int co = i & 3, r = i >> 2;
        const int ci = r % Cin, tap = r / Cin;
        sw[i] = co < Cout ? w[((size_t)co * Cin + ci) * 9 + tap] : 0.f;
    }
    __syncthreads();
    const int b = blockIdx.y;
    const int HW = H * W;
    const int pp = blockIdx.x * blockDim.x + threadIdx.x;
    if (pp >= HW) return;
    const int h = pp / W, ww = pp - h * W;
    float acc[4] = {0, 0, 0, 0};
    for (int tap = 0; tap < 9; ++tap) {
        const int gh = h + tap / 3 - 1;
        int gw = ww + tap % 3 - 1;
        if (gh < 0 || gh >= H) continue;
        if (gw < 0) { if (ring) gw += W; else continue; }
        else if (gw >= W) { if (ring) gw -= W; else continue; }
        const TIn* ap = a + ((size_t)b * HW + (size_t)gh * W + gw) * Cin;
        const float4* wt = sw4 + tap * Cin;
        for (int ci = 0; ci < Cin; ci += 8) {
            float v[8];
            ld8(ap + ci, v);
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const float4 wv = wt[ci + e];
                acc[0] = fmaf(v[e], wv.x, acc[0]);
                acc[1] = fmaf(v[e], wv.y, acc[1]);
                acc[2] = fmaf(v[e], wv.z, acc[2]);
                acc[3] = fmaf(v[e], wv.w, acc[3]);
            }
        }
    }
    for (int co = 0; co < Cout; ++co) pred[((size_t)b * Cout + co) * HW + pp] = acc[co] + bias[co];
}

// out_conv, row-tiled: one block = 128 consecutive pixels of one image row.  The three halo rows are staged through
// shared memory in 16-channel chunks with COALESCED loads (4 lanes x 16 B per pixel; a thread-per-pixel walk over the
// 256-byte pixel rows costs 32 L1 line look-ups per load instruction and ran at 0.1 ms on B200).  Stage 1: thread =
// one staged pixel, reduces its channels to the 3 x Cout partial sums of the filter row it feeds; stage 2 adds the 9
// partial sums of each output pixel.
constexpr int OC_PIX = 128, OC_THREADS = 96, OC_CH = 16, OC_PITCH = 20, OC_PPL = 5;   // warp = staged row, lane = 5 pixels

__global__ void __launch_bounds__(OC_THREADS) out_conv_rows_kernel(const float* __restrict__ a, const float* __restrict__ w,
                                                                   const float* __restrict__ bias, float* __restrict__ pred,
                                                                   int H, int W, int Cin, int Cout, int ring) {
    extern __shared__ __align__(16) float osm[];
    pdl_launch_dependents();
    pdl_wait();
    float* sw = osm;                                   // [3 dy][Cin][12]  (dx*4 + co, zero padded)
    float* sx = osm + 3 * Cin * 12;                    // [3 rows][130 px][OC_PITCH] channel chunk; later [3][130][12] sums
    for (int i = threadIdx.x; i < 3 * Cin * 12; i += OC_THREADS) {
        const int q = i % 12, ci = (i / 12) % Cin, dy = i / (12 * Cin);
        const int dx = q >> 2, co = q & 3;
        sw[i] = co < Cout ? w[((size_t)co * Cin + ci) * 9 + dy * 3 + dx] : 0.f;
    }
    const int wt = blockIdx.x, h = blockIdx.y, b = blockIdx.z;
    const int w0 = wt * OC_PIX;
    constexpr int NPX = OC_PIX + 2;
    // Stage 1: warp r owns staged row r, lane l owns pixels l, l + 32, ... (consecutive lanes -> consecutive pixels:
    // conflict-free 16-byte reads at the 20-word pitch).  Every weight float4 read from shared memory feeds 5 pixels:
    // the one-pixel-per-thread version was bound by those LDS.128 (ncu: l1tex 71 %, 101 us on B200).
    const int ir = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float acc[OC_PPL][12];
#pragma unroll
    for (int i = 0; i < OC_PPL; ++i)
#pragma unroll
        for (int q = 0; q < 12; ++q) acc[i][q] = 0.f;
    for (int c0 = 0; c0 < Cin; c0 += OC_CH) {
        __syncthreads();                               // previous chunk consumed (and, first time, sw complete)
        for (int i = threadIdx.x; i < 3 * NPX * (OC_CH / 4); i += OC_THREADS) {
            const int f = i % (OC_CH / 4), pj = i / (OC_CH / 4);
            const int r = pj / NPX, j = pj - r * NPX;
            const int gh = h + r - 1;
            int gw = w0 + j - 1;
            bool ok = gh >= 0 && gh < H;
            if (gw < 0) { if (ring) gw += W; else ok = false; }
            else if (gw >= W) { if (ring) gw -= W; else ok = false; }
            const float4 v = ok ? ld4(a + ((size_t)(b * H + gh) * W + gw) * Cin + c0 + 4 * f) : make_float4(0, 0, 0, 0);
            *reinterpret_cast<float4*>(sx + (size_t)pj * OC_PITCH + 4 * f) = v;
        }
        __syncthreads();
        const float* wr = sw + ((size_t)ir * Cin + c0) * 12;
#pragma unroll
        for (int f = 0; f < OC_CH / 4; ++f) {
            float4 xv[OC_PPL];
#pragma unroll
            for (int i = 0; i < OC_PPL; ++i) {
                const int j = lane + 32 * i;
                xv[i] = j < NPX ? *reinterpret_cast<const float4*>(sx + (size_t)(ir * NPX + j) * OC_PITCH + 4 * f)
                                : make_float4(0, 0, 0, 0);
            }
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const float4 w0v = *reinterpret_cast<const float4*>(wr + (4 * f + e) * 12);
                const float4 w1v = *reinterpret_cast<const float4*>(wr + (4 * f + e) * 12 + 4);
                const float4 w2v = *reinterpret_cast<const float4*>(wr + (4 * f + e) * 12 + 8);
#pragma unroll
                for (int i = 0; i < OC_PPL; ++i) {
                    const float x = e == 0 ? xv[i].x : e == 1 ? xv[i].y : e == 2 ? xv[i].z : xv[i].w;
                    acc[i][0] = fmaf(x, w0v.x, acc[i][0]); acc[i][1] = fmaf(x, w0v.y, acc[i][1]);
                    acc[i][2] = fmaf(x, w0v.z, acc[i][2]); acc[i][3] = fmaf(x, w0v.w, acc[i][3]);
                    acc[i][4] = fmaf(x, w1v.x, acc[i][4]); acc[i][5] = fmaf(x, w1v.y, acc[i][5]);
                    acc[i][6] = fmaf(x, w1v.z, acc[i][6]); acc[i][7] = fmaf(x, w1v.w, acc[i][7]);
                    acc[i][8] = fmaf(x, w2v.x, acc[i][8]); acc[i][9] = fmaf(x, w2v.y, acc[i][9]);
                    acc[i][10] = fmaf(x, w2v.z, acc[i][10]); acc[i][11] = fmaf(x, w2v.w, acc[i][11]);
                }
            }
        }
    }
    __syncthreads();                                   // all chunk reads done: reuse sx for the partial sums
    float* st = sx;                                    // [3 rows][130 px][12]
#pragma unroll
    for (int i = 0; i < OC_PPL; ++i) {
        const int j = lane + 32 * i;
        if (j < NPX) {
#pragma unroll
            for (int q = 0; q < 12; ++q) st[(size_t)(ir * NPX + j) * 12 + q] = acc[i][q];
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < OC_PIX * Cout; i += OC_THREADS) {
        const int co = i / OC_PIX, px = i - co * OC_PIX;     // consecutive threads -> consecutive pixels (NCHW store)
        float o = bias[co];
#pragma unroll
        for (int r = 0; r < 3; ++r)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) o += st[(size_t)(r * NPX + px + dx) * 12 + dx * 4 + co];
        pred[((size_t)(b * Cout + co) * H + h) * W + w0 + px] = o;
    }
}

// out_conv for Cout = 2 (the denoiser's eps prediction), third version: per-pixel projection + 3x3 gather.
//   g[p][tap][co] = sum_c a[p][c] * w[co][c][tap]       (18 dot products of length Cin per INPUT pixel, read once, coalesced
//                                                         256-byte rows, weights broadcast from shared memory)
//   out[p][co]    = bias[co] + sum_tap g[p + delta(tap)][tap][co]        (3x3 gather inside the block's shared-memory tile)
// One block = OG_ROWS x 128 output pixels (+ a one-pixel halo of g: 1.27x recompute).  The row-staged kernel above was
// instruction-bound (ncu: 53 M warp instructions for 19 M useful FMAs, 15 warps / SM, 125 us); here every weight float4
// feeds two pixels and there is no per-chunk staging loop.
constexpr int OG_ROWS = 8, OG_PX = 128, OG_THREADS = 256, OG_N = 18, OG_PITCH = 19;

__global__ void __launch_bounds__(OG_THREADS) out_conv_gather_kernel(const float* __restrict__ a, const float* __restrict__ w,
                                                                     const float* __restrict__ bias, float* __restrict__ pred,
                                                                     int H, int W, int Cin, int ring) {
    extern __shared__ __align__(16) float gsm[];
    pdl_launch_dependents();
    pdl_wait();
    const int Q = Cin / 4;
    float* sw = gsm;                              // [Q][18][4]: for channel quad q and n = tap * 2 + co the 4 weights
    float* sg = gsm + (size_t)Q * OG_N * 4;       // [(OG_ROWS + 2)][OG_PX + 2][OG_PITCH]
    for (int i = threadIdx.x; i < Q * OG_N * 4; i += OG_THREADS) {
        const int e = i & 3, n = (i >> 2) % OG_N, q = i / (4 * OG_N);
        const int tap = n >> 1, co = n & 1;
        sw[i] = w[((size_t)co * Cin + q * 4 + e) * 9 + tap];
    }
    __syncthreads();
    const int w0 = blockIdx.x * OG_PX, h0 = blockIdx.y * OG_ROWS, b = blockIdx.z;
    constexpr int NPX = OG_PX + 2, NR = OG_ROWS + 2, NP = NPX * NR;
    // ---- phase 1: g for the tile + halo; a thread projects two pixels per pass (weights loaded once for both) ----
    for (int i0 = threadIdx.x; i0 < NP; i0 += 2 * OG_THREADS) {
        const float* src[2];
        bool ok[2];
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = i0 + u * OG_THREADS;
            const int r = i / NPX, j = i - r * NPX;
            const int gh = h0 + r - 1;
            int gw = w0 + j - 1;
            ok[u] = i < NP && gh >= 0 && gh < H;
            if (gw < 0) { if (ring) gw += W; else ok[u] = false; }
            else if (gw >= W) { if (ring) gw -= W; else ok[u] = false; }
            src[u] = ok[u] ? a + ((size_t)(b * H + gh) * W + gw) * Cin : a;
        }
        float acc[2][OG_N];
#pragma unroll
        for (int u = 0; u < 2; ++u)
#pragma unroll
            for (int n = 0; n < OG_N; ++n) acc[u][n] = 0.f;
        for (int q = 0; q < Q; ++q) {
            const float4 x0 = ok[0] ? ld4(src[0] + 4 * q) : make_float4(0, 0, 0, 0);
            const float4 x1 = ok[1] ? ld4(src[1] + 4 * q) : make_float4(0, 0, 0, 0);
            const float4* wq = reinterpret_cast<const float4*>(sw) + q * OG_N;
#pragma unroll
            for (int n = 0; n < OG_N; ++n) {
                const float4 wv = wq[n];
                acc[0][n] = fmaf(x0.x, wv.x, fmaf(x0.y, wv.y, fmaf(x0.z, wv.z, fmaf(x0.w, wv.w, acc[0][n]))));
                acc[1][n] = fmaf(x1.x, wv.x, fmaf(x1.y, wv.y, fmaf(x1.z, wv.z, fmaf(x1.w, wv.w, acc[1][n]))));
            }
        }
#pragma unroll
        for (int u = 0; u < 2; ++u) {
            const int i = i0 + u * OG_THREADS;
            if (i < NP) {
#pragma unroll
                for (int n = 0; n < OG_N; ++n) sg[(size_t)i * OG_PITCH + n] = acc[u][n];
            }
        }
    }
    __syncthreads();
    // ---- phase 2: 3x3 gather; consecutive threads -> consecutive pixels of one (co, row): coalesced NCHW stores ----
    for (int o = threadIdx.x; o < 2 * OG_ROWS * OG_PX; o += OG_THREADS) {
        const int px = o % OG_PX, r = (o / OG_PX) % OG_ROWS, co = o / (OG_PX * OG_ROWS);
        if (h0 + r >= H) continue;
        float v = bias[co];
#pragma unroll
        for (int dy = 0; dy < 3; ++dy)
#pragma unroll
            for (int dx = 0; dx < 3; ++dx) v += sg[(size_t)((r + dy) * NPX + px + dx) * OG_PITCH + (dy * 3 + dx) * 2 + co];
        pred[((size_t)(b * 2 + co) * H + h0 + r) * W + w0 + px] = v;
    }
}

// ---------------------------------------------------------------------------------------------------------
// sampler update (continuous_time.py:205-231)
// ---------------------------------------------------------------------------------------------------------
// x_t and x_s MAY ALIAS (the captured denoiser step updates the sampler state in place): no __restrict__ on them; every
// thread reads element gi before it writes element gi.
__global__ void sampler_update_kernel(const float* x_t, const float* __restrict__ pred,
                                      const float* __restrict__ noise, const float* __restrict__ coef,
                                      float* x_s, int n, int mode, int objective, float clip) {
    pdl_launch_dependents();
    pdl_wait();
    const int b = blockIdx.y;
    const float a_t = coef[b * 8 + 0], s_t = coef[b * 8 + 1], a_s = coef[b * 8 + 2], s_s = coef[b * 8 + 3];
    const float c1 = coef[b * 8 + 4], c2 = coef[b * 8 + 5], cc = coef[b * 8 + 6];
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const size_t gi = (size_t)b * n + i;
        const float xt = x_t[gi], pr = pred[gi];
        float x0;
        if (objective == 0) x0 = (xt - s_t * pr) / a_t;
        else if (objective == 1) x0 = a_t * xt - s_t * pr;
        else x0 = pr;
        if (clip > 0.f) x0 = fminf(fmaxf(x0, -clip), clip);
        float out;
        if (mode == 0) {
            const float eps = (xt - a_t * x0) / s_t;
            out = a_s * x0 + c2 * eps;
            if (noise) out += c1 * noise[gi];
        } else {
            const float mean = a_s * (xt * (1.f - cc) / a_t + cc * x0);
            out = mean + s_s * sqrtf(cc) * noise[gi];
        }
        x_s[gi] = out;
    }
}

// ---------------------------------------------------------------------------------------------------------
// sampler coefficients: log-SNR schedules + DDIM/DDPM step coefficients of a whole batch in ONE launch
// (continuous_time.py:14-63 and :200-231 evaluate them with ~25 tiny elementwise kernels per p_step)
// ---------------------------------------------------------------------------------------------------------
struct ScheduleParams {
    int kind;             // 0 linear, 1 cosine, 2 cosine_shifted, 3 cosine_interpolated
    float t_min, t_span;  // cosine family: atan(exp(-logsnr_max/2)), t_max - t_min
    float shift_lo, shift_hi;   // 2 log(noise_d / image_d) for noise_d_low / noise_d_high
};

__device__ __forceinline__ float log_snr_f(const ScheduleParams& sp, float t) {
    if (sp.kind == 0) return -logf(fmaxf(expm1f(__fadd_rn(1e-4f, __fmul_rn(10.f, __fmul_rn(t, t)))), 1e-20f));
    const float base = -2.f * logf(fmaxf(tanf(__fadd_rn(sp.t_min, __fmul_rn(t, sp.t_span))), 1e-20f));
    if (sp.kind == 1) return base;
    if (sp.kind == 2) return base + sp.shift_lo;
    return t * (base + sp.shift_lo) + (1.f - t) * (base + sp.shift_hi);
}

__device__ __forceinline__ float sigmoid_f(float x) { return 1.f / (1.f + expf(-x)); }

__global__ void sampler_coef_kernel(const float* __restrict__ step_t, const float* __restrict__ step_s, ScheduleParams sp,
                                    float eta, float* __restrict__ log_snr_t, float* __restrict__ coef, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const float lt = log_snr_f(sp, step_t[b]), ls = log_snr_f(sp, step_s[b]);
    const float a_t = sqrtf(sigmoid_f(lt)), s_t = sqrtf(sigmoid_f(-lt));
    const float a_s = sqrtf(sigmoid_f(ls)), s_s = sqrtf(sigmoid_f(-ls));
    const float c1 = eta * s_s / s_t * sqrtf(1.f - (a_t * a_t) / (a_s * a_s));
    const float c2 = sqrtf(1.f - a_s * a_s - c1 * c1);
    const float cc = -expm1f(lt - ls);
    log_snr_t[b] = lt;
    float* o = coef + (size_t)b * 8;
    o[0] = a_t; o[1] = s_t; o[2] = a_s; o[3] = s_s; o[4] = c1; o[5] = c2; o[6] = cc; o[7] = 0.f;
}

}  // namespace b200

using namespace b200;

extern "C" int b200_sampler_coefficients(const float* step_t, const float* step_s, int schedule, float t_min, float t_span,
                                         float shift_lo, float shift_hi, float ddim_eta, float* log_snr_t, float* coef,
                                         int B, void* stream) {
    B200_CHECK_ARG(step_t && step_s && log_snr_t && coef && B > 0 && schedule >= 0 && schedule <= 3);
    ScheduleParams sp{schedule, t_min, t_span, shift_lo, shift_hi};
    sampler_coef_kernel<<<cdiv(B, 128), 128, 0, (cudaStream_t)stream>>>(step_t, step_s, sp, ddim_eta, log_snr_t, coef, B);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

extern "C" int b200_version(void) { return 100; }
extern "C" const char* b200_last_error(void) { return g_err; }
extern "C" int b200_device_check(int dev) {
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) {
        set_error("cudaGetDeviceProperties(%d) failed", dev);
        return B200_E_CUDA;
    }
    if (prop.major != 10) {
        set_error("device %d is sm_%d%d, libb200lidar needs sm_100", dev, prop.major, prop.minor);
        return B200_E_ARCH;
    }
    return prop.multiProcessorCount;
}

static bool c4_ok(int C) { return C % 4 == 0 && C / 4 <= 256 && 256 % (C / 4) == 0; }

static int num_sms_cached() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

extern "C" int b200_gn_act_f16(const float* x0, int C0, const float* x1, int C1, const double* stats0,
                               const double* stats1, const float* gamma, const float* beta, const float* ada,
                               int ada_stride, int groups, float eps, int silu, void* y, void* y_raw, int parts, int B,
                               int H, int W, void* stream) {
    B200_CHECK_ARG(parts >= 1 && parts <= 3);
    B200_CHECK_ARG(W % OTW == 0);   // the conv operand is organised in 128-pixel tiles
    B200_CHECK_ARG(x0 && y && C0 > 0 && C0 % 8 == 0 && C1 % 8 == 0 && (C1 == 0 || x1));
    const int C = C0 + C1;
    B200_CHECK_ARG(parts != 3 || C % 32 == 0);   // fp8 plane: whole warps per 32-channel item
    B200_CHECK_ARG(C <= GN_MAX_C);
    if (stats0) {
        B200_CHECK_ARG(groups > 0 && groups <= 64 && C % groups == 0);
        B200_CHECK_ARG(C1 == 0 || stats1);
        B200_CHECK_ARG((gamma == nullptr) == (beta == nullptr));
    } else {
        B200_CHECK_ARG(!gamma && !ada);
    }
    // two decompositions (measured on B200, profiles/r01_gn_act_variants.txt): few large blocks of 256 pixels x all
    // channels win on big tensors; channel-split blocks in one resident wave win when there are only a few tiles per
    // sample (4x128 levels: 12 us instead of 18 us).  B200_GN_V1=0/1 forces one of them.
    static int forced = -2, bps = 8;
    if (forced == -2) {
        const char* e = getenv("B200_GN_V1");
        forced = e ? (e[0] == '1' ? 1 : 0) : -1;
        if (const char* t = getenv("B200_GN_BPS")) bps = atoi(t) > 0 ? atoi(t) : 8;
    }
    const int variant = forced >= 0 ? forced : ((long long)H * W * B > 8192 ? 1 : 0);
    if (variant == 1) {
        const int HW = H * W;
        int ppb_log2 = 8;                                    // 256 pixels per block, never more than one image row
        while ((1 << ppb_log2) > W) --ppb_log2;
        while (ppb_log2 > 5 && (long long)(HW >> ppb_log2) * B < 2 * 148) --ppb_log2;
        dim3 grid(HW >> ppb_log2, B);
        const size_t plane = (size_t)B * H * (W / OTW) * ((C0 + C1) / 8) * OPX * 8;
#define B200_GN_V1(P_, R_)                                                                                              \
    launch_pdl(gn_act_kernel_v1<P_, R_>, grid, dim3(256), (size_t)24 * C, (cudaStream_t)stream, x0, C0, x1, C1, stats0,    \
               stats1, gamma, beta, ada, ada_stride, groups, eps, silu, (__half*)y, (__half*)y_raw, plane, HW, W, ppb_log2)
        if (y_raw) {
            if (parts == 1) B200_GN_V1(1, true); else if (parts == 2) B200_GN_V1(2, true); else B200_GN_V1(3, true);
        } else {
            if (parts == 1) B200_GN_V1(1, false); else if (parts == 2) B200_GN_V1(2, false); else B200_GN_V1(3, false);
        }
#undef B200_GN_V1
        B200_CHECK_LAUNCH();
        return B200_OK;
    }
    // grid = one resident wave: channel blocks of CB (a multiple of 32) channels x strided tile sets, ~4 blocks per SM
    const int tiles = H * (W / OTW);
    const int target = bps * num_sms_cached();
    int CB = C >= 128 ? 128 : ((C + 31) / 32) * 32;
    while (CB > 32 && (long long)cdiv(C, CB) * tiles * B < target) CB >>= 1;
    const int n_cb = cdiv(C, CB);
    int tile_blocks = target / (n_cb * B);
    if (tile_blocks < 1) tile_blocks = 1;
    if (tile_blocks > tiles) tile_blocks = tiles;
    const int cpg = stats0 ? C / groups : 1;
    const size_t smem = (size_t)8 * CB + (size_t)16 * (CB + 2 * cpg);
    B200_CHECK_ARG(smem <= 48 * 1024);
    dim3 grid(n_cb * tile_blocks, B);
    launch_pdl(gn_act_kernel, grid, dim3(256), smem, (cudaStream_t)stream, x0, C0, x1, C1, stats0, stats1, gamma, beta, ada,
               ada_stride, groups, eps, silu, (__half*)y, (__half*)y_raw,
               (size_t)B * H * (W / OTW) * ((C0 + C1) / 8) * OPX * 8, parts, H, W, CB, tile_blocks);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

extern "C" int b200_gn_act_f32(const float* x, const double* stats, const float* gamma, const float* beta, int groups,
                               float eps, int silu, float* y, int B, int HW, int C, void* stream) {
    B200_CHECK_ARG(x && stats && y && C % 4 == 0 && C <= GN_MAX_C);
    B200_CHECK_ARG(groups > 0 && groups <= 64 && C % groups == 0);
    B200_CHECK_ARG((gamma == nullptr) == (beta == nullptr));
    int ppb = 256;
    while (ppb > 32 && (long long)cdiv(HW, ppb) * B < 2 * 148) ppb >>= 1;
    dim3 grid(cdiv(HW, ppb), B);
    launch_pdl(gn_act_f32_kernel, grid, dim3(256), (size_t)24 * C, (cudaStream_t)stream, x, C, stats, gamma, beta, groups,
               eps, silu, y, (double*)nullptr, HW, ppb);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

extern "C" int b200_channel_stats(const float* x, double* stats, int B, int HW, int C, void* stream) {
    B200_CHECK_ARG(x && stats && c4_ok(C));
    dim3 grid(cdiv(HW, ST_PIX_PER_BLOCK), B);
    channel_stats_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, stats, HW, C);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

extern "C" int b200_fir_resample(const float* x, float* y, double* stats, int B, int H, int W, int C, int up, int ring,
                                 void* stream) {
    B200_CHECK_ARG(x && y && c4_ok(C));
    B200_CHECK_ARG(up || (H % 2 == 0 && W % 2 == 0));
    const int npo = up ? 4 * H * W : (H / 2) * (W / 2);
    int ppb = ST_PIX_PER_BLOCK;
    const int pmin = 256 / (C / 4) > 8 ? 256 / (C / 4) : 8;   // at least one pixel per thread row
    while (ppb > pmin && (long long)cdiv(npo, ppb) * B < 4 * 148) ppb >>= 1;
    if (!up && (W / 2) % 2 == 0) {      // two outputs per thread
        const int npairs = npo / 2;
        int pb = ST_PIX_PER_BLOCK / 2;
        const int pbmin = 256 / (C / 4) > 4 ? 256 / (C / 4) : 4;
        while (pb > pbmin && (long long)cdiv(npairs, pb) * B < 4 * 148) pb >>= 1;
        launch_pdl(fir_down_pair_kernel, dim3(cdiv(npairs, pb), B), dim3(256), 0, (cudaStream_t)stream, x, y, stats, H, W, C,
                   ring, pb);
        B200_CHECK_LAUNCH();
        return B200_OK;
    }
    dim3 grid(cdiv(npo, ppb), B);
    if (up) launch_pdl(fir_kernel<true>, grid, dim3(256), 0, (cudaStream_t)stream, x, y, stats, H, W, C, ring, ppb);
    else launch_pdl(fir_kernel<false>, grid, dim3(256), 0, (cudaStream_t)stream, x, y, stats, H, W, C, ring, ppb);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

extern "C" int b200_fir_up_operand(const float* x, void* y, int parts, int B, int H, int W, int C, int ring, void* stream) {
    B200_CHECK_ARG(x && y && parts >= 1 && parts <= 3);
    B200_CHECK_ARG(C % 8 == 0 && (parts != 3 || C % 32 == 0) && (2 * W) % OTW == 0);
    dim3 grid(2 * H * (2 * W / OTW), B);
    launch_pdl(fir_up_operand_kernel, grid, dim3(256), 0, (cudaStream_t)stream, x, (__half*)y,
               (size_t)B * 2 * H * (2 * W / OTW) * (C / 8) * OPX * 8, parts, H, W, C, ring);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

extern "C" int b200_time_embed(const float* t, const float* w1, const float* b1, const float* w2, const float* b2,
                               const float* temb_add, const float* wp, const float* bp, float* temb, float* ada, int B,
                               int Cs, int E, int P, void* stream) {
    B200_CHECK_ARG(t && w1 && b1 && w2 && b2 && temb);
    B200_CHECK_ARG(B > 0 && B <= 65535 && Cs % 2 == 0 && Cs >= 4);
    launch_pdl(temb_kernel, dim3(B, cdiv(E, TEMB_EO)), dim3(256), (Cs + E) * sizeof(float), (cudaStream_t)stream, t, w1, b1,
               w2, b2, temb_add, temb, Cs, E);
    B200_CHECK_LAUNCH();
    if (P > 0) {
        B200_CHECK_ARG(wp && bp && ada);
        B200_CHECK_ARG((size_t)ADA_MAX_B * E * sizeof(float) <= 48 * 1024);
        for (int b0 = 0; b0 < B; b0 += ADA_MAX_B) {      // the kernel keeps one accumulator per sample in registers
            const int nb = B - b0 < ADA_MAX_B ? B - b0 : ADA_MAX_B;
            launch_pdl(ada_proj_kernel, dim3(cdiv(P, ADA_ROWS_PER_BLOCK)), dim3(256), (size_t)nb * E * sizeof(float),
                       (cudaStream_t)stream, (const float*)temb + (size_t)b0 * E, wp, bp, ada + (size_t)b0 * P, nb, E, P);
            B200_CHECK_LAUNCH();
        }
    }
    return B200_OK;
}

extern "C" int b200_in_conv(const float* x, const float* w, const float* cst, int cst_batched, float* out,
                            double* stats, int B, int H, int W, int Cx, int Cout, int ring, void* stream) {
    B200_CHECK_ARG(x && w && cst && out);
    B200_CHECK_ARG(Cx >= 1 && Cx <= 4 && Cout <= 256 && c4_ok(Cout));
    if (W % IC_PIX == 0 && Cout <= 128) {
        int rpb = 4;                                   // rows per block, as long as the grid still fills the GPU twice
        while (rpb > 1 && (long long)(W / IC_PIX) * cdiv(H, rpb) * B < 2 * 148) rpb >>= 1;
        dim3 grid(W / IC_PIX, cdiv(H, rpb), B);
        launch_pdl(in_conv_rows_kernel, grid, dim3(256), 0, (cudaStream_t)stream, x, w, cst, cst_batched, out, stats, H, W,
                   Cx, Cout, ring, rpb);
        B200_CHECK_LAUNCH();
        return B200_OK;
    }
    dim3 grid(cdiv(H * W, ST_PIX_PER_BLOCK), B);
    launch_pdl(in_conv_kernel, grid, dim3(256), 0, (cudaStream_t)stream, x, w, cst, cst_batched, out, stats, H, W, Cx, Cout,
               ring);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

extern "C" int b200_conv_direct_f32(const float* x, const float* w, const float* bias, float* out, int B, int H, int W,
                                    int Cin, int Cout, int k, int ring, void* stream) {
    B200_CHECK_ARG(x && w && out && (k == 1 || k == 3));
    const long long total = (long long)B * H * W * Cout;
    const int blocks = (int)((total + 255) / 256 > 65535 ? 65535 : (total + 255) / 256);
    conv_direct_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(x, w, bias, out, B, H, W, Cin, Cout, k, ring);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

extern "C" int b200_out_conv(const void* a, int a_is_f16, const float* w, const float* bias, float* pred, int B, int H,
                             int W, int Cin, int Cout, int ring, void* stream) {
    B200_CHECK_ARG(a && w && bias && pred && Cout >= 1 && Cout <= 4);
    B200_CHECK_ARG(Cin % 8 == 0);
    static int og_ok = -1;     // B200_OUT_CONV=rows selects the previous kernel (A/B timing)
    if (og_ok < 0) {
        const char* e = getenv("B200_OUT_CONV");
        og_ok = (e && e[0] == 'r') ? 0 : 1;
    }
    if (og_ok && !a_is_f16 && Cout == 2 && W % OG_PX == 0 && Cin % 4 == 0) {
        const size_t sm = ((size_t)(Cin / 4) * OG_N * 4 + (size_t)(OG_ROWS + 2) * (OG_PX + 2) * OG_PITCH) * sizeof(float);
        if (sm <= 200 * 1024) {
            static bool attr_set = false;
            if (!attr_set) {
                if (cudaFuncSetAttribute(out_conv_gather_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024) !=
                    cudaSuccess) {
                    set_error("out_conv: cudaFuncSetAttribute failed");
                    return B200_E_CUDA;
                }
                attr_set = true;
            }
            dim3 grid(W / OG_PX, cdiv(H, OG_ROWS), B);
            launch_pdl(out_conv_gather_kernel, grid, dim3(OG_THREADS), sm, (cudaStream_t)stream, (const float*)a, w, bias, pred,
                       H, W, Cin, ring);
            B200_CHECK_LAUNCH();
            return B200_OK;
        }
    }
    if (!a_is_f16 && W % OC_PIX == 0 && Cin % OC_CH == 0) {
        const size_t sm = ((size_t)3 * Cin * 12 + 3 * (OC_PIX + 2) * OC_PITCH) * sizeof(float);
        if (sm <= 48 * 1024) {
            dim3 grid(W / OC_PIX, H, B);
            launch_pdl(out_conv_rows_kernel, grid, dim3(OC_THREADS), sm, (cudaStream_t)stream, (const float*)a, w, bias, pred,
                       H, W, Cin, Cout, ring);
            B200_CHECK_LAUNCH();
            return B200_OK;
        }
    }
    const size_t smem = (size_t)4 * 9 * Cin * sizeof(float);
    B200_CHECK_ARG(smem <= 48 * 1024);
    dim3 grid(cdiv(H * W, 128), B);
    if (a_is_f16)
        out_conv_kernel<__half><<<grid, 128, smem, (cudaStream_t)stream>>>((const __half*)a, w, bias, pred, H, W, Cin,
                                                                          Cout, ring);
    else
        out_conv_kernel<float><<<grid, 128, smem, (cudaStream_t)stream>>>((const float*)a, w, bias, pred, H, W, Cin,
                                                                         Cout, ring);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

extern "C" int b200_sampler_update(const float* x_t, const float* pred, const float* noise, const float* coef,
                                   float* x_s, int B, int n_per_sample, int mode, int objective, float clip,
                                   void* stream) {
    B200_CHECK_ARG(x_t && pred && coef && x_s && (mode == 0 || (mode == 1 && noise)));
    dim3 grid(cdiv(n_per_sample, 256 * 4), B);
    launch_pdl(sampler_update_kernel, grid, dim3(256), 0, (cudaStream_t)stream, x_t, pred, noise, coef, x_s, n_per_sample,
               mode, objective, clip);
    B200_CHECK_LAUNCH();
    return B200_OK;
}
