// K2 (round-1 form): softmax(q k^T * scale) v per (batch, head) with the score rows of a 16-query tile kept in
// shared memory (never materialised in HBM, unlike the reference's [B*h, T, S] fp32 tensor).
// replaces nn.MultiheadAttention's core (efficient_unet.py:39-53) and QKVAttentionLegacy
// (layout_unet_v1.py:488-505).  fp32 CUDA-core math; output fp16 (operand of the out-projection GEMM).
#include "common.cuh"

namespace b200 {

constexpr int ATT_QT = 16;        // queries per CTA
constexpr int ATT_THREADS = 128;

struct AttnParams {
    const float *q, *k, *v;
    __half* out;
    size_t lo_off;
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
        // slab-major conv operand: token t = (h, w) of an image of width Wimg; [b][h][ldo/8][w][8]
        const int tq = q0 + qi;
        const int hh = tq / p.Wimg, ww = tq - hh * p.Wimg;
        const int ch = head * p.dv + dc * dper;
        __half* o = p.out + ((((size_t)b * (p.Tq / p.Wimg) + hh) * (p.ldo / 8) + ch / 8) * p.Wimg + ww) * 8 + (ch & 7);
        for (int e = 0; e < dper; ++e) {
            const float val = acc[e] * inv;
            const __half hi = __float2half_rn(val);
            o[e] = hi;
            if (p.lo_off) o[p.lo_off + e] = __float2half_rn(val - __half2float(hi));
        }
    }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_attention(const float* q, int ldq, int qoff, const float* k, int ldk, int koff, const float* v,
                              int ldv, int voff, void* out, int ldo, int out_w, int parts, int B, int heads, int Tq, int Tk,
                              int dqk, int dv, float scale, void* stream) {
    B200_CHECK_ARG(parts == 1 || parts == 2);
    B200_CHECK_ARG(out_w > 0 && Tq % out_w == 0 && ldo % 8 == 0);
    B200_CHECK_ARG(q && k && v && out);
    B200_CHECK_ARG(dqk % 4 == 0 && (dv == 32 || dv == 64));
    B200_CHECK_ARG(ldq % 4 == 0 && ldk % 4 == 0 && ldv % 4 == 0 && qoff % 4 == 0 && koff % 4 == 0 && voff % 4 == 0);
    AttnParams p{q, k, v, (__half*)out, parts == 2 ? (size_t)B * Tq * ldo : 0, ldq, qoff, ldk, koff, ldv, voff, ldo, out_w, heads, Tq, Tk, dqk, dv, scale};
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
    size_t lo_off;
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
        __half* o = p.out + ((((size_t)b * (p.T / p.Wimg) + hh) * (p.C / 8) + ch / 8) * p.Wimg + ww) * 8 + (ch & 7);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const float val = acc[e] * inv;
            const __half hi = __float2half_rn(val);
            o[e] = hi;
            if (p.lo_off) o[p.lo_off + e] = __float2half_rn(val - __half2float(hi));
        }
    }
}

}  // namespace b200

extern "C" int b200_attention_oa(const float* qkv, const float* pos_p, const float* kl, const float* pos_l,
                                 const float* vl, void* out, int out_w, int parts, int B, int C, int heads, int T,
                                 int L2, float scale2, void* stream) {
    B200_CHECK_ARG(qkv && pos_p && kl && pos_l && vl && out);
    B200_CHECK_ARG(parts == 1 || parts == 2);
    B200_CHECK_ARG(heads > 0 && C % heads == 0 && C / heads == 32);   // num_head_channels = 32 in every config
    B200_CHECK_ARG(out_w > 0 && T % out_w == 0 && L2 >= 0);
    OAParams p{qkv, pos_p, kl, pos_l, vl, (__half*)out, parts == 2 ? (size_t)B * T * C : 0, C, heads, T, L2, C / heads,
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
