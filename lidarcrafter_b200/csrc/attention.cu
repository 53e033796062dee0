// K2 (round-1 form): softmax(q k^T * scale) v per (batch, head) with the score rows of a 16-query tile kept in
// shared memory (never materialised in HBM, unlike the reference's [B*h, T, S] fp32 tensor).
// replaces nn.MultiheadAttention's core (efficient_unet.py:39-53) and QKVAttentionLegacy
// (layout_unet_v1.py:488-505).  fp32 CUDA-core math; output fp16 (operand of the out-projection GEMM).
#include <stdlib.h>

#include "common.cuh"

namespace b200 {

constexpr int ATT_QT = 16;        // queries per CTA
constexpr int ATT_THREADS = 128;

struct AttnParams {
    const float *q, *k, *v;
    __half* out;
    size_t lo_off;   // fp16 elements per operand plane
    int parts;
    int ldq, qoff, ldk, koff, ldv, voff, ldo, Wimg;
    int heads, Tq, Tk, dqk, dv;
    float scale;
};

__global__ void __launch_bounds__(ATT_THREADS) attention_kernel(const AttnParams p) {
    extern __shared__ float sm[];
    const int ldS = p.Tk + 1;
    float* sq = sm;                       // [16][dqk]
    float* sS = sm + ATT_QT * p.dqk;      // [16][Tk+1]
    float* sInv = sS + ATT_QT * ldS;      // [16]
    const int q0 = blockIdx.x * ATT_QT, head = blockIdx.y, b = blockIdx.z;
    const int nq = min(ATT_QT, p.Tq - q0);
    const int tid = threadIdx.x;

    for (int i = tid; i < ATT_QT * p.dqk; i += ATT_THREADS) {
        const int qi = i / p.dqk, d = i - qi * p.dqk;
        sq[i] = qi < nq ? p.q[((size_t)b * p.Tq + q0 + qi) * p.ldq + p.qoff + head * p.dqk + d] * p.scale : 0.f;
    }
    __syncthreads();

    // phase 1: scores
    for (int j = tid; j < p.Tk; j += ATT_THREADS) {
        const float* kr = p.k + ((size_t)b * p.Tk + j) * p.ldk + p.koff + head * p.dqk;
        float acc[ATT_QT];
#pragma unroll
        for (int qi = 0; qi < ATT_QT; ++qi) acc[qi] = 0.f;
        for (int d = 0; d < p.dqk; d += 4) {
            const float4 kv = *reinterpret_cast<const float4*>(kr + d);
#pragma unroll
            for (int qi = 0; qi < ATT_QT; ++qi) {
                const float4 qv = *reinterpret_cast<const float4*>(sq + qi * p.dqk + d);
                acc[qi] = fmaf(qv.x, kv.x, acc[qi]);
                acc[qi] = fmaf(qv.y, kv.y, acc[qi]);
                acc[qi] = fmaf(qv.z, kv.z, acc[qi]);
                acc[qi] = fmaf(qv.w, kv.w, acc[qi]);
            }
        }
#pragma unroll
        for (int qi = 0; qi < ATT_QT; ++qi) sS[qi * ldS + j] = acc[qi];
    }
    __syncthreads();

    // phase 2: row softmax (4 warps x 4 rows)
    const int warp = tid >> 5, lane = tid & 31;
    for (int qi = warp; qi < ATT_QT; qi += ATT_THREADS / 32) {
        float* row = sS + qi * ldS;
        float m = -INFINITY;
        for (int j = lane; j < p.Tk; j += 32) m = fmaxf(m, row[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = 0.f;
        for (int j = lane; j < p.Tk; j += 32) {
            const float e = __expf(row[j] - m);
            row[j] = e;
            s += e;
        }
        s = warp_sum(s);
        if (lane == 0) sInv[qi] = 1.f / s;
    }
    __syncthreads();

    // phase 3: out = P V ; thread -> (query qi = tid/8, dv slice (tid%8) * dv/8)
    const int qi = tid >> 3, dc = tid & 7;
    const int dper = p.dv / 8;  // 4 or 8
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    const float* vb = p.v + (size_t)b * p.Tk * p.ldv + p.voff + head * p.dv + dc * dper;
    const float* prow = sS + qi * ldS;
    for (int j = 0; j < p.Tk; ++j) {
        const float pj = prow[j];
        const float* vr = vb + (size_t)j * p.ldv;
        const float4 v0 = *reinterpret_cast<const float4*>(vr);
        acc[0] = fmaf(pj, v0.x, acc[0]); acc[1] = fmaf(pj, v0.y, acc[1]);
        acc[2] = fmaf(pj, v0.z, acc[2]); acc[3] = fmaf(pj, v0.w, acc[3]);
        if (dper == 8) {
            const float4 v1 = *reinterpret_cast<const float4*>(vr + 4);
            acc[4] = fmaf(pj, v1.x, acc[4]); acc[5] = fmaf(pj, v1.y, acc[5]);
            acc[6] = fmaf(pj, v1.z, acc[6]); acc[7] = fmaf(pj, v1.w, acc[7]);
        }
    }
    if (qi < nq) {
        const float inv = sInv[qi];
        // conv operand (common.cuh): token t = (h, w) of an image of width Wimg
        const int tq = q0 + qi;
        const int hh = tq / p.Wimg, ww = tq - hh * p.Wimg;
        const int ch = head * p.dv + dc * dper;
        for (int e = 0; e < dper; ++e)
            store_operand_elem(p.out, p.lo_off, p.parts, (size_t)b * (p.Tq / p.Wimg) + hh, p.ldo, p.Wimg, ww, ch + e,
                               acc[e] * inv);
    }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_attention(const float* q, int ldq, int qoff, const float* k, int ldk, int koff, const float* v,
                              int ldv, int voff, void* out, int ldo, int out_w, int parts, int B, int heads, int Tq, int Tk,
                              int dqk, int dv, float scale, void* stream) {
    B200_CHECK_ARG(parts >= 1 && parts <= 3);
    B200_CHECK_ARG(out_w > 0 && out_w % OTW == 0 && Tq % out_w == 0 && ldo % 8 == 0);
    B200_CHECK_ARG(q && k && v && out);
    B200_CHECK_ARG(dqk % 4 == 0 && (dv == 32 || dv == 64));
    B200_CHECK_ARG(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && qoff % 4 == 0 && koff % 4 == 0 && voff % 4 == 0);
    AttnParams p{q, k, v, (__half*)out, (size_t)B * Tq / out_w * (out_w / OTW) * (ldo / 8) * OPX * 8, parts, ldq, qoff, ldk, koff, ldv, voff, ldo, out_w, heads, Tq, Tk, dqk, dv, scale};
    const size_t smem = ((size_t)ATT_QT * dqk + (size_t)ATT_QT * (Tk + 1) + ATT_QT) * sizeof(float);
    B200_CHECK_ARG(smem <= 200 * 1024);
    static size_t smem_set = 0;
    if (smem > 48 * 1024 && smem > smem_set) {
        cudaError_t e = cudaFuncSetAttribute(attention_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) {
            set_error("attention: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
            return B200_E_CUDA;
        }
        smem_set = 200 * 1024;
    }
    dim3 grid(cdiv(Tq, ATT_QT), heads, B);
    attention_kernel<<<grid, ATT_THREADS, smem, (cudaStream_t)stream>>>(p);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

// =========================================================================================================
// ObjectAwareCrossAttention core (layout_unet_v1.py:416-505): image tokens attend to image tokens AND layout
// objects; query/key = [content ; positional] (2d channels per head), value = content only.
//   score(t, s)     = scale2 * ( q_c(t).k_c(s)  + pos_p(t).pos_p(s) )        s < T   (image keys)
//   score(t, T + j) = scale2 * ( q_c(t).k_l(j)  + pos_p(t).pos_l(j) )        j < L2  (layout keys)
//   out(t)          = softmax_s(score) . [ v_c ; v_l ]
// qkv fp32 [B,T,3C] (q | k | v, head-major channels), pos_p fp32 [B,T,C], kl/pos_l/vl fp32 [B,L2,C].
// Same 16-query-tile / smem-score-row structure as attention_kernel (scores never touch HBM).
// =========================================================================================================
namespace b200 {

struct OAParams {
    const float *qkv, *pos_p, *kl, *pos_l, *vl;
    __half* out;
    size_t lo_off;   // fp16 elements per operand plane
    int parts;
    int C, heads, T, L2, d, Wimg;
    float scale2;
};

__global__ void __launch_bounds__(ATT_THREADS) attention_oa_kernel(const OAParams p) {
    extern __shared__ float sm[];
    const int S = p.T + p.L2;
    const int ldS = S + 1;
    const int d = p.d, d2 = 2 * p.d;
    float* sq = sm;                    // [16][2d]  (content | positional), pre-scaled
    float* sS = sm + ATT_QT * d2;      // [16][S+1]
    float* sInv = sS + ATT_QT * ldS;   // [16]
    const int q0 = blockIdx.x * ATT_QT, head = blockIdx.y, b = blockIdx.z;
    const int nq = min(ATT_QT, p.T - q0);
    const int tid = threadIdx.x;
    const int C3 = 3 * p.C;

    for (int i = tid; i < ATT_QT * d2; i += ATT_THREADS) {
        const int qi = i / d2, c = i - qi * d2;
        float v = 0.f;
        if (qi < nq) {
            const size_t tok = (size_t)b * p.T + q0 + qi;
            v = c < d ? p.qkv[tok * C3 + head * d + c] : p.pos_p[tok * p.C + head * d + (c - d)];
        }
        sq[i] = v * p.scale2;
    }
    __syncthreads();

    for (int j = tid; j < S; j += ATT_THREADS) {
        const float *k1, *k2;
        if (j < p.T) {
            const size_t tok = (size_t)b * p.T + j;
            k1 = p.qkv + tok * C3 + p.C + head * d;
            k2 = p.pos_p + tok * p.C + head * d;
        } else {
            const size_t tok = (size_t)b * p.L2 + (j - p.T);
            k1 = p.kl + tok * p.C + head * d;
            k2 = p.pos_l + tok * p.C + head * d;
        }
        float acc[ATT_QT];
#pragma unroll
        for (int qi = 0; qi < ATT_QT; ++qi) acc[qi] = 0.f;
        for (int c = 0; c < d; c += 4) {
            const float4 kv = *reinterpret_cast<const float4*>(k1 + c);
            const float4 pv = *reinterpret_cast<const float4*>(k2 + c);
#pragma unroll
            for (int qi = 0; qi < ATT_QT; ++qi) {
                const float4 qv = *reinterpret_cast<const float4*>(sq + qi * d2 + c);
                const float4 qp = *reinterpret_cast<const float4*>(sq + qi * d2 + d + c);
                float a = acc[qi];
                a = fmaf(qv.x, kv.x, a); a = fmaf(qv.y, kv.y, a); a = fmaf(qv.z, kv.z, a); a = fmaf(qv.w, kv.w, a);
                a = fmaf(qp.x, pv.x, a); a = fmaf(qp.y, pv.y, a); a = fmaf(qp.z, pv.z, a); a = fmaf(qp.w, pv.w, a);
                acc[qi] = a;
            }
        }
#pragma unroll
        for (int qi = 0; qi < ATT_QT; ++qi) sS[qi * ldS + j] = acc[qi];
    }
    __syncthreads();

    const int warp = tid >> 5, lane = tid & 31;
    for (int qi = warp; qi < ATT_QT; qi += ATT_THREADS / 32) {
        float* row = sS + qi * ldS;
        float m = -INFINITY;
        for (int j = lane; j < S; j += 32) m = fmaxf(m, row[j]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
        float s = 0.f;
        for (int j = lane; j < S; j += 32) {
            const float e = __expf(row[j] - m);
            row[j] = e;
            s += e;
        }
        s = warp_sum(s);
        if (lane == 0) sInv[qi] = 1.f / s;
    }
    __syncthreads();

    // out = P [v ; vl]: thread -> (query tid/8, 4 value channels (tid%8)*4), d == 32
    const int qi = tid >> 3, dc = tid & 7;
    float acc[4] = {0, 0, 0, 0};
    const float* prow = sS + qi * ldS;
    const float* vb = p.qkv + (size_t)b * p.T * C3 + 2 * p.C + head * d + dc * 4;
    for (int j = 0; j < p.T; ++j) {
        const float pj = prow[j];
        const float4 v0 = *reinterpret_cast<const float4*>(vb + (size_t)j * C3);
        acc[0] = fmaf(pj, v0.x, acc[0]); acc[1] = fmaf(pj, v0.y, acc[1]);
        acc[2] = fmaf(pj, v0.z, acc[2]); acc[3] = fmaf(pj, v0.w, acc[3]);
    }
    const float* vx = p.vl + (size_t)b * p.L2 * p.C + head * d + dc * 4;
    for (int j = 0; j < p.L2; ++j) {
        const float pj = prow[p.T + j];
        const float4 v0 = *reinterpret_cast<const float4*>(vx + (size_t)j * p.C);
        acc[0] = fmaf(pj, v0.x, acc[0]); acc[1] = fmaf(pj, v0.y, acc[1]);
        acc[2] = fmaf(pj, v0.z, acc[2]); acc[3] = fmaf(pj, v0.w, acc[3]);
    }
    if (qi < nq) {
        const float inv = sInv[qi];
        const int tq = q0 + qi;
        const int hh = tq / p.Wimg, ww = tq - hh * p.Wimg;
        const int ch = head * d + dc * 4;
#pragma unroll
        for (int e = 0; e < 4; ++e)
            store_operand_elem(p.out, p.lo_off, p.parts, (size_t)b * (p.T / p.Wimg) + hh, p.C, p.Wimg, ww, ch + e,
                               acc[e] * inv);
    }
}

}  // namespace b200

extern "C" int b200_attention_oa(const float* qkv, const float* pos_p, const float* kl, const float* pos_l,
                                 const float* vl, void* out, int out_w, int parts, int B, int C, int heads, int T,
                                 int L2, float scale2, void* stream) {
    B200_CHECK_ARG(qkv && pos_p && kl && pos_l && vl && out);
    B200_CHECK_ARG(parts >= 1 && parts <= 3);
    B200_CHECK_ARG(heads > 0 && C % heads == 0 && C / heads == 32);   // num_head_channels = 32 in every config
    B200_CHECK_ARG(out_w > 0 && out_w % OTW == 0 && T % out_w == 0 && L2 >= 0);
    OAParams p{qkv, pos_p, kl, pos_l, vl, (__half*)out, (size_t)B * T / out_w * (out_w / OTW) * (C / 8) * OPX * 8, parts, C, heads, T, L2, C / heads,
               out_w, scale2};
    const size_t smem = ((size_t)ATT_QT * 2 * p.d + (size_t)ATT_QT * (T + L2 + 1) + ATT_QT) * sizeof(float);
    B200_CHECK_ARG(smem <= 200 * 1024);
    static bool attr = false;
    if (!attr) {
        cudaError_t e = cudaFuncSetAttribute(attention_oa_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
        if (e != cudaSuccess) {
            set_error("attention_oa: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
            return B200_E_CUDA;
        }
        attr = true;
    }
    dim3 grid(cdiv(T, ATT_QT), heads, B);
    attention_oa_kernel<<<grid, ATT_THREADS, smem, (cudaStream_t)stream>>>(p);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

// =========================================================================================================
// Flash-style fp32 attention (online softmax, 64-query x 64-key register-tiled blocks): replaces the two
// score-row kernels above on the hot path.  One kernel, two operand loaders:
//   MHA : q/k/v = column slices of the fused qkv tensor (efficient_unet.py:39-53), d = dv = 32 or 64
//   OA  : q = [q_c ; pos_p], k = [[k_c ; pos_p] | [k_l ; pos_l]], v = [v_c | v_l], d = 32 + 32, dv = 32
//         (layout_unet_v1.py:488-505)
// Scores never leave the SM; all math fp32 (the tolerance budget is spent nowhere here).
// =========================================================================================================
namespace b200 {

constexpr int FA_BQ = 64, FA_BK = 64, FA_THREADS = 256, FA_PAD = 68;

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
};

template <int DQ, int DV>
__global__ void __launch_bounds__(FA_THREADS) flash_attn_kernel(const FAParams p) {
    extern __shared__ float sm[];
    pdl_launch_dependents();
    pdl_wait();
    float* sQ = sm;                       // [DQ][FA_PAD]   (d-major: 4 consecutive queries = one LDS.128)
    float* sK = sQ + DQ * FA_PAD;         // [DQ][FA_PAD]
    float* sV = sK + DQ * FA_PAD;         // [FA_BK][DV]
    float* sP = sV + FA_BK * DV;          // [FA_BK][FA_PAD] (key-major)
    const int q0 = blockIdx.x * FA_BQ, head = blockIdx.y, b = blockIdx.z;
    const int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
    constexpr int DPT = DV / 16;          // value dims per thread

    // ---- load the query tile (pre-scaled), transposed ----
    for (int i = tid; i < FA_BQ * (DQ / 4); i += FA_THREADS) {
        const int qi = i / (DQ / 4), c4 = (i - qi * (DQ / 4)) * 4;
        const int tq = q0 + qi;
        float4 v = make_float4(0, 0, 0, 0);
        if (tq < p.T) {
            const size_t tok = (size_t)b * p.T + tq;
            v = c4 < p.d1 ? *reinterpret_cast<const float4*>(p.q1 + tok * p.ldq1 + head * p.d1 + c4)
                          : *reinterpret_cast<const float4*>(p.q2 + tok * p.ldq2 + head * p.d2 + (c4 - p.d1));
        }
        sQ[(c4 + 0) * FA_PAD + qi] = v.x * p.scale; sQ[(c4 + 1) * FA_PAD + qi] = v.y * p.scale;
        sQ[(c4 + 2) * FA_PAD + qi] = v.z * p.scale; sQ[(c4 + 3) * FA_PAD + qi] = v.w * p.scale;
    }
    float m_run[4], l_run[4], o[4][DPT];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        m_run[i] = -INFINITY; l_run[i] = 0.f;
#pragma unroll
        for (int j = 0; j < DPT; ++j) o[i][j] = 0.f;
    }
    const int S = p.T + p.Tx;
    const int ntiles = (S + FA_BK - 1) / FA_BK;
    for (int kt = 0; kt < ntiles; ++kt) {
        const int k0 = kt * FA_BK;
        __syncthreads();   // previous tile's sK / sV / sP fully consumed (also orders the sQ stores on iteration 0)
        // ---- K tile (transposed) and V tile ----
        for (int i = tid; i < FA_BK * (DQ / 4); i += FA_THREADS) {
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
            sK[(c4 + 0) * FA_PAD + kj] = v.x; sK[(c4 + 1) * FA_PAD + kj] = v.y;
            sK[(c4 + 2) * FA_PAD + kj] = v.z; sK[(c4 + 3) * FA_PAD + kj] = v.w;
        }
        for (int i = tid; i < FA_BK * (DV / 4); i += FA_THREADS) {
            const int kj = i / (DV / 4), c4 = (i - kj * (DV / 4)) * 4;
            const int tk = k0 + kj;
            float4 v = make_float4(0, 0, 0, 0);
            if (tk < p.T) v = *reinterpret_cast<const float4*>(p.v + ((size_t)b * p.T + tk) * p.ldv + head * DV + c4);
            else if (tk < S) v = *reinterpret_cast<const float4*>(p.xv + ((size_t)b * p.Tx + (tk - p.T)) * p.ldx + head * DV + c4);
            *reinterpret_cast<float4*>(sV + kj * DV + c4) = v;
        }
        __syncthreads();
        // ---- S = Q K^T for this thread's 4 queries x 4 keys ----
        float s[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) s[i][j] = 0.f;
#pragma unroll 8
        for (int d = 0; d < DQ; ++d) {
            const float4 qv = *reinterpret_cast<const float4*>(sQ + d * FA_PAD + ty * 4);
            const float4 kv = *reinterpret_cast<const float4*>(sK + d * FA_PAD + tx * 4);
            const float qa[4] = {qv.x, qv.y, qv.z, qv.w}, ka[4] = {kv.x, kv.y, kv.z, kv.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) s[i][j] = fmaf(qa[i], ka[j], s[i][j]);
        }
        // mask keys past the end
#pragma unroll
        for (int j = 0; j < 4; ++j)
            if (k0 + tx * 4 + j >= S)
#pragma unroll
                for (int i = 0; i < 4; ++i) s[i][j] = -INFINITY;
        // ---- online softmax: row max over the 16 threads (tx) sharing the same queries ----
        float alpha[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            float mx = fmaxf(fmaxf(s[i][0], s[i][1]), fmaxf(s[i][2], s[i][3]));
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            const float m_new = fmaxf(m_run[i], mx);
            alpha[i] = __expf(m_run[i] - m_new);      // exp(-inf) = 0 on the first tile
            m_run[i] = m_new;
            float rs = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float e = __expf(s[i][j] - m_new);
                s[i][j] = e;
                rs += e;
            }
#pragma unroll
            for (int off = 8; off > 0; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
            l_run[i] = l_run[i] * alpha[i] + rs;
#pragma unroll
            for (int j = 0; j < DPT; ++j) o[i][j] *= alpha[i];
        }
        // P -> smem (key-major so that 4 queries are contiguous)
#pragma unroll
        for (int j = 0; j < 4; ++j)
            *reinterpret_cast<float4*>(sP + (tx * 4 + j) * FA_PAD + ty * 4) = make_float4(s[0][j], s[1][j], s[2][j], s[3][j]);
        __syncthreads();
        // ---- O += P V : this thread's 4 queries x DPT value dims (dims tx*DPT ..) ----
#pragma unroll 8
        for (int k = 0; k < FA_BK; ++k) {
            const float4 pv = *reinterpret_cast<const float4*>(sP + k * FA_PAD + ty * 4);
            const float pa[4] = {pv.x, pv.y, pv.z, pv.w};
            float va[DPT];
            if constexpr (DPT == 4) {
                const float4 vv = *reinterpret_cast<const float4*>(sV + k * DV + tx * 4);
                va[0] = vv.x; va[1] = vv.y; va[2] = vv.z; va[3] = vv.w;
            } else {
                const float2 vv = *reinterpret_cast<const float2*>(sV + k * DV + tx * 2);
                va[0] = vv.x; va[1] = vv.y;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < DPT; ++j) o[i][j] = fmaf(pa[i], va[j], o[i][j]);
        }
    }
    // ---- normalise and store (conv operand layout, common.cuh) ----
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int tq = q0 + ty * 4 + i;
        if (tq >= p.T) continue;
        const float inv = 1.f / l_run[i];
        const int hh = tq / p.Wimg, ww = tq - hh * p.Wimg;
        const int ch = head * DV + tx * DPT;
#pragma unroll
        for (int j = 0; j < DPT; ++j)
            store_operand_elem(p.out, p.lo_off, p.parts, (size_t)b * (p.T / p.Wimg) + hh, p.C, p.Wimg, ww, ch + j,
                               o[i][j] * inv);
    }
}


// ---------------------------------------------------------------------------------------------------------
// Tensor-core version of the same contract (the one the plans launch).  64 queries x 64 keys per step, 4 warps, one
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

template <int DQ, int DV>
static int launch_fa(const FAParams& p, int B, int heads, cudaStream_t st) {
    static int ffma = -1;       // B200_FA_FFMA=1: the CUDA-core kernel above (A/B timing, cross-check)
    if (ffma < 0) {
        const char* e = getenv("B200_FA_FFMA");
        ffma = (e && e[0] == '1') ? 1 : 0;
    }
    if (!ffma) {
        const size_t smem = ((size_t)2 * FM_BK * (DQ + 8) + (size_t)2 * DV * (FM_BK + 8)) * sizeof(__half);
        dim3 grid(cdiv(p.T, FM_BQ), heads, B);
        launch_pdl(flash_attn_mma_kernel<DQ, DV>, grid, dim3(FM_THREADS), smem, st, p);
        B200_CHECK_LAUNCH();
        return B200_OK;
    }
    const size_t smem = ((size_t)2 * DQ * FA_PAD + FA_BK * DV + FA_BK * FA_PAD) * sizeof(float);
    static bool attr = false;
    if (!attr && smem > 48 * 1024) {
        cudaError_t e = cudaFuncSetAttribute(flash_attn_kernel<DQ, DV>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) {
            set_error("flash_attn: cudaFuncSetAttribute failed: %s", cudaGetErrorString(e));
            return B200_E_CUDA;
        }
        attr = true;
    }
    dim3 grid(cdiv(p.T, FA_BQ), heads, B);
    launch_pdl(flash_attn_kernel<DQ, DV>, grid, dim3(FA_THREADS), smem, st, p);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

}  // namespace b200

extern "C" int b200_flash_attention(const float* qkv, int E, void* out, int out_w, int parts, int B, int heads, int T,
                                    float scale, void* stream) {
    // self-attention on the fused in-projection output qkv fp32 [B,T,3E] (q | k | v, head-major)
    B200_CHECK_ARG(qkv && out && (parts >= 1 && parts <= 3) && heads > 0 && E % heads == 0);
    B200_CHECK_ARG(out_w > 0 && out_w % OTW == 0 && T % out_w == 0);
    const int d = E / heads;
    FAParams p{qkv, nullptr, qkv + E, nullptr, qkv + 2 * E, nullptr, nullptr, nullptr, 3 * E, 0, 3 * E, 0, 3 * E, 0,
               d, 0, d, T, 0, (__half*)out, (size_t)B * T / out_w * (out_w / OTW) * (E / 8) * OPX * 8, parts, E, out_w, scale};
    if (d == 64) return launch_fa<64, 64>(p, B, heads, (cudaStream_t)stream);
    if (d == 32) return launch_fa<32, 32>(p, B, heads, (cudaStream_t)stream);
    set_error("flash_attention: head dim %d not supported (32 or 64)", d);
    return B200_E_ARG;
}

extern "C" int b200_flash_attention_oa(const float* qkv, const float* pos_p, const float* kl, const float* pos_l,
                                       const float* vl, void* out, int out_w, int parts, int B, int C, int heads, int T,
                                       int L2, float scale2, void* stream) {
    B200_CHECK_ARG(qkv && pos_p && kl && pos_l && vl && out && (parts >= 1 && parts <= 3));
    B200_CHECK_ARG(heads > 0 && C % heads == 0 && C / heads == 32 && out_w > 0 && out_w % OTW == 0 && T % out_w == 0 && L2 >= 0);
    FAParams p{qkv, pos_p, qkv + C, pos_p, qkv + 2 * C, kl, pos_l, vl, 3 * C, C, 3 * C, C, 3 * C, C,
               32, 32, 32, T, L2, (__half*)out, (size_t)B * T / out_w * (out_w / OTW) * (C / 8) * OPX * 8, parts, C, out_w, scale2};
    return launch_fa<64, 32>(p, B, heads, (cudaStream_t)stream);
}
