// Shared helpers for libb200lidar (sm_100a only): error plumbing + thin PTX wrappers for
// mbarrier / cp.async / cp.async.bulk (TMA engine) / tcgen05 (tensor cores + TMEM).
#pragma once
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/b200lidar.h"

namespace b200 {

void set_error(const char* fmt, ...);

#define B200_CHECK_ARG(cond)                                                        \
    do {                                                                            \
        if (!(cond)) {                                                              \
            b200::set_error("%s:%d: bad argument: %s", __FILE__, __LINE__, #cond);  \
            return B200_E_ARG;                                                      \
        }                                                                           \
    } while (0)

#define B200_CHECK_LAUNCH()                                                                        \
    do {                                                                                           \
        cudaError_t e__ = cudaGetLastError();                                                      \
        if (e__ != cudaSuccess) {                                                                  \
            b200::set_error("%s:%d: CUDA error: %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return B200_E_CUDA;                                                                    \
        }                                                                                          \
    } while (0)

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }

// elementwise.cu: out = ((sum_s part[s]) * w_inv + bias + res) * scale (+ per-channel statistics) -- the conv epilogue on the
// partial sums of the K slices of b200_conv_tc_splitk
int launch_splitk_reduce(const float* part, int splits, size_t stride, const float* bias, const float* res, float w_inv,
                         float scale, float* out, double* stats, int B, int HW, int Cout, void* stream);

// Programmatic dependent launch (PDL): every kernel of the denoiser step is launched with
// cudaLaunchAttributeProgrammaticStreamSerialization, calls pdl_launch_dependents() first thing and pdl_wait() before
// its first global-memory access.  The next kernel's CTAs are then scheduled (and run their prologue: barrier init,
// TMEM allocation, coefficient setup) while the tail of the previous kernel is still draining.  Opt-in with B200_PDL=1 (measured neutral-to-slightly-negative inside CUDA graphs).
bool pdl_enabled();
bool pdl_enabled_conv();

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl_if(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem,
                                        cudaStream_t st, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

template <typename... KArgs, typename... Args>
static inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                     Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// ------------------------------------------------------------------------------------------------
// device-side PTX wrappers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// ---- programmatic dependent launch ----
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- mbarrier ----
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// try_wait WITH a suspend-time hint: the hardware parks the warp until the phase completes (or the hint, in ns, expires).
// Without the hint the wait returns after a short system-defined time and the poll loop re-issues at a high rate: measured
// in the fused conv (transform warps doing FP32 / MUFU work next to ~10 polling warps), the pollers took so many issue slots
// that the transform ran 2.8x slower than with the barriers already complete (profiles/r02_fused_front_ablation.txt).
constexpr uint32_t MBAR_SUSPEND_NS = 10000;
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity), "r"(MBAR_SUSPEND_NS)
        : "memory");
    return ok != 0;
}
// Bounded wait (~2^18 x 10 us): a protocol bug becomes a trap (launch error) instead of a hung GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 18)) {
            printf("b200lidar: mbarrier timeout (block %d,%d,%d thread %d bar 0x%x parity %u)\n", blockIdx.x,
                   blockIdx.y, blockIdx.z, threadIdx.x, bar, parity);
            __trap();
        }
    }
}

// same bound, no printf: a CALL inside a register-heavy loop makes ptxas spill everything that is live across it
__device__ __forceinline__ void mbar_wait_quiet(uint32_t bar, uint32_t parity) {
    uint32_t spins = 0;
    while (!mbar_try_wait(bar, parity)) {
        if (++spins > (1u << 18)) __trap();
    }
}

// ---- named barriers between a subset of warps (count = participating threads, arrivers + waiters) ----
__device__ __forceinline__ void named_bar_sync(uint32_t id, uint32_t count) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory");
}
__device__ __forceinline__ void named_bar_arrive(uint32_t id, uint32_t count) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory");
}

// ---- cp.async (LDGSTS) ----
__device__ __forceinline__ void cp_async_16(uint32_t dst, const void* src, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}
// generic-proxy writes (st.shared / cp.async) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// ---- cp.async.bulk (TMA engine, 1-D) ----
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

// ---- tcgen05 ----
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T, kind::f16 (fp16/bf16 operands, fp32 accumulate), issued by ONE thread
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// same, kind::f8f6f4 (here: E4M3 x E4M3, K = 32 per instruction, fp32 accumulate)
__device__ __forceinline__ void tc_mma_f8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// one lane of a converged warp (the warp keeps executing uniformly; tcgen05.mma / commit operands stay in uniform registers)
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}
// Descriptors split in 32-bit halves: the HIGH word (SBO = 128 B, version 1) is the same constant for every operand of
// the conv kernel, the LOW word = (address >> 4) | (LBO >> 4) << 16, so stepping to another tap / row / K group is ONE
// 32-bit add of a compile-time constant on a per-stage base (a single issuing thread is latency-bound: every
// instruction between two MMAs counts).
constexpr uint32_t DESC_HI_SBO128 = (128u >> 4) | (1u << 14);
__device__ __forceinline__ uint32_t desc_lo(uint32_t saddr, uint32_t lbo_bytes) {
    return ((saddr >> 4) & 0x3FFFu) | ((lbo_bytes >> 4) << 16);
}
__device__ __forceinline__ void tc_mma_f16_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(DESC_HI_SBO128), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_mma_f8_lh(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc,
                                             uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "setp.ne.b32 p, %5, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(DESC_HI_SBO128), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
//   element (row r, k) lives at start + (r/8)*sbo + (r%8)*16 + (k/8)*lbo + (k%8)*2   [fp16]
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
    return d;                // base_offset 0, lbo_mode 0, layout_type 0 (SWIZZLE_NONE)
}
// instruction descriptor: D fp32, A/B fp16 (kind::f16) or E4M3 (kind::f8f6f4: format code 0 as well), K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc_f16(int M, int N) {
    return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// 32 lanes x 32 columns of fp32 accumulators -> 32 registers per thread (thread i = TMEM lane base+i)
__device__ __forceinline__ void tmem_ld_32x32(uint32_t taddr, float* v) {
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
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// x * sigmoid(x) = x / (1 + 2^(-x log2 e)) with the two MUFU approximations (~2 ulp) in their flush-to-zero form: 3 FP + 2
// MUFU instructions (the non-ftz __expf / __fdividef forms carry a denormal fix-up per call); identical results wherever
// the intermediate values are normal, and 1 + denormal == 1 either way.  The result is rounded to fp16 hi + lo right after.
__device__ __forceinline__ float silu_f(float x) {
    float e, r;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x * -1.4426950408889634f));
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
    return x * r;
}

// ---- "fp16 + fp8 correction" operand encoding (conv precision mode parts = 3) ----
//   x = hi16 + lo,  lo ~ 2^-11 |x|:   plane 0 keeps hi16 (fp16);  plane 1 keeps, per 16-channel chunk, two 16-byte
//   units per pixel: L8 = e4m3(lo * 2^11) and A8 = e4m3(x).  The conv then computes
//   hi16 x w16  (kind::f16)  +  [L8 | A8] x [e4m3(w 2^-11) | e4m3(w - w16)]  (ONE kind::f8f6f4 MMA, K = 32).
constexpr float F8_LO_SCALE = 2048.f;
__device__ __forceinline__ uint32_t f8x4(float a, float b, float c, float d) {
    const uint32_t lo = __nv_cvt_float2_to_fp8x2(make_float2(a, b), __NV_SATFINITE, __NV_E4M3);
    const uint32_t hi = __nv_cvt_float2_to_fp8x2(make_float2(c, d), __NV_SATFINITE, __NV_E4M3);
    return lo | (hi << 16);
}
__device__ __forceinline__ uint8_t f8x1(float a) {
    return (uint8_t)__nv_cvt_float_to_fp8(a, __NV_SATFINITE, __NV_E4M3);
}
__device__ __forceinline__ float f8_to_float(uint8_t v) {
    const __half_raw h = __nv_cvt_fp8_to_halfraw(v, __NV_E4M3);
    return __half2float(__half(h));
}
// ---- conv operand layout ("tile-major slabs") ----
//   [plane][b][h][W/128][C/8][130][8] fp16-sized elements: for one image row, one 128-pixel tile and one 8-channel
//   group, 130 pixels at a 16-byte pitch = the tile's 128 pixels preceded / followed by ONE halo pixel (the ring
//   neighbours (w0-1) mod W and (w0+128) mod W).  This is exactly the tcgen05 no-swizzle K-major shared-memory image
//   of the tile INCLUDING the 3x3 halo, and the KG slabs of a K chunk are contiguous, so the conv stages one
//   (row, plane, chunk) with ONE cp.async.bulk (the TMA engine is request-rate bound: 3 copies per slab -- body plus
//   two 16-byte halo pixels -- cost 18% of the kernel on B200).  Producers write the two duplicated halo pixels.
constexpr int OPX = 130;    // pixels per slab
constexpr int OTW = 128;    // tile width
// 16-byte unit index of pixel position `pos` (0..129) of slab (bh, wt, cg)
__host__ __device__ __forceinline__ size_t operand_unit(size_t bh, int WT, int CG, int wt, int cg, int pos) {
    return ((bh * WT + wt) * CG + cg) * OPX + pos;
}
// home position of pixel ww and (if it is a tile-border pixel) its halo duplicate in the neighbouring tile
struct OperandPos {
    int wt, pos, wt2, pos2;   // wt2 < 0: no duplicate
};
__device__ __forceinline__ OperandPos operand_pos(int ww, int WT) {
    OperandPos r;
    r.wt = ww >> 7;
    const int x = ww & (OTW - 1);
    r.pos = x + 1;
    r.wt2 = -1;
    r.pos2 = 0;
    if (x == 0) { r.wt2 = r.wt == 0 ? WT - 1 : r.wt - 1; r.pos2 = OPX - 1; }
    else if (x == OTW - 1) { r.wt2 = r.wt == WT - 1 ? 0 : r.wt + 1; r.pos2 = 0; }
    return r;
}
// scalar store of channel `ch` of pixel (bh = b*H + h, ww) into a conv operand (parts = 1, 2 or 3);
// plane_elems = fp16-sized elements per plane = B*H*(W/128)*(C/8)*130*8
__device__ __forceinline__ void store_operand_elem(__half* out, size_t plane_elems, int parts, size_t bh, int C, int W,
                                                   int ww, int ch, float val) {
    const int WT = W / OTW;
    const OperandPos op = operand_pos(ww, WT);
    const __half hi = __float2half_rn(val);
    const __half lo = __float2half_rn(val - __half2float(hi));
    const uint8_t l8 = parts == 3 ? f8x1((val - __half2float(hi)) * F8_LO_SCALE) : 0;
    const uint8_t a8 = parts == 3 ? f8x1(val) : 0;
#pragma unroll
    for (int rep = 0; rep < 2; ++rep) {
        const int wt = rep == 0 ? op.wt : op.wt2, pos = rep == 0 ? op.pos : op.pos2;
        if (wt < 0) break;
        const size_t u = operand_unit(bh, WT, C / 8, wt, ch / 8, pos);
        out[u * 8 + (ch & 7)] = hi;
        if (parts == 2) {
            out[plane_elems + u * 8 + (ch & 7)] = lo;
        } else if (parts == 3) {
            // plane 1: slab (2*chunk) holds L8, slab (2*chunk + 1) holds A8 of the 16-channel chunk
            uint8_t* p1 = reinterpret_cast<uint8_t*>(out + plane_elems);
            const size_t u8 = operand_unit(bh, WT, C / 8, wt, (ch / 16) * 2, pos);
            p1[u8 * 16 + (ch & 15)] = l8;
            p1[(u8 + OPX) * 16 + (ch & 15)] = a8;
        }
    }
}

// second operand plane of one lane's 8 channels (v = fp32 values, h = their fp16 roundings), as one 16-byte unit:
//   parts == 2: lo = fp16(x - hi)  (error-compensation term of the fp16x3 split)
//   parts == 3: per 16-channel chunk the even 8-channel slab position holds L8 = e4m3((x - hi) * 2^11) of all 16
//               channels and the odd one A8 = e4m3(x) (see common.cuh); the two lanes of a chunk swap halves.
__device__ __forceinline__ uint4 plane1_value(const float* v, const __half2* h, int parts, int lane) {
    float lo[8];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float2 hf = __half22float2(h[e]);
        lo[2 * e] = v[2 * e] - hf.x;
        lo[2 * e + 1] = v[2 * e + 1] - hf.y;
    }
    if (parts == 2) {
        __half2 l[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) l[e] = __floats2half2_rn(lo[2 * e], lo[2 * e + 1]);
        return *reinterpret_cast<const uint4*>(l);
    }
    uint2 l8, a8;
    l8.x = f8x4(lo[0] * F8_LO_SCALE, lo[1] * F8_LO_SCALE, lo[2] * F8_LO_SCALE, lo[3] * F8_LO_SCALE);
    l8.y = f8x4(lo[4] * F8_LO_SCALE, lo[5] * F8_LO_SCALE, lo[6] * F8_LO_SCALE, lo[7] * F8_LO_SCALE);
    a8.x = f8x4(v[0], v[1], v[2], v[3]);
    a8.y = f8x4(v[4], v[5], v[6], v[7]);
    const bool odd = lane & 1;
    const uint2 send = odd ? l8 : a8;
    uint2 recv;
    recv.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
    recv.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
    return odd ? make_uint4(recv.x, recv.y, a8.x, a8.y) : make_uint4(l8.x, l8.y, recv.x, recv.y);
}

// GroupNorm-apply (+SiLU) of one lane's 8 channels and store as conv operand unit(s): v <- act(v * a + b); writes the
// fp16 hi unit, the plane-1 unit (parts >= 2) and the halo duplicates.  All 32 lanes of a warp must call it together
// when parts == 3 (plane1_value shuffles between the two lanes of a 16-channel chunk).
__device__ __forceinline__ void gn_apply_store(float* v, const float* av, const float* bv, int silu, int parts, __half* y,
                                               size_t lo_off, size_t oi, size_t oi2, bool dup, int lane) {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
        float t = fmaf(v[e], av[e], bv[e]);
        if (silu) t = silu_f(t);
        v[e] = t;
    }
    __half2 h[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) h[e] = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
    const uint4 hv = *reinterpret_cast<const uint4*>(h);
    *reinterpret_cast<uint4*>(y + oi) = hv;
    if (dup) *reinterpret_cast<uint4*>(y + oi2) = hv;
    if (parts >= 2) {   // parts 2: lo = fp16(x - fp32(hi)); parts 3: e4m3 pair plane
        const uint4 pv = plane1_value(v, h, parts, lane);
        *reinterpret_cast<uint4*>(y + lo_off + oi) = pv;
        if (dup) *reinterpret_cast<uint4*>(y + lo_off + oi2) = pv;
    }
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

}  // namespace b200
