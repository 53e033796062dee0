// Probe for the NEXT conv design: tcgen05.mma with the A operand in TENSOR MEMORY.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_ts_probe mma_ts_probe.cu && ./mma_ts_probe
// NOT RUN YET (written at the end of round 1 / session 2 when the GPU budget was spent; it compiles for sm_100a).
//
// Why: conv_tc_kernel reads BOTH operands of every MMA from shared memory (A 4 KB + B 32*N bytes per M=128, K=16 MMA).
// The tensor pipe fetches operands at 128 B/clk, so N = 64 MMAs run at 48 instead of 32 cycles and N = 128 MMAs sit exactly
// at the limit, where every TMA fill and epilogue transpose costs tensor time (profiles/r01_mma_probe.txt,
// profiles/r01s2_conv_ablation.txt).  With the WEIGHTS as the TMEM-resident A operand (M = output channels, staged once
// per K chunk with tcgen05.cp) and the PIXELS as the B operand (N = 128..256 pixels, the tap shift is still just a start
// address), an MMA reads 32*N bytes of shared memory in 0.5*N cycles = 64 B/clk, and the accumulator comes out as
// [channel lane][pixel column], i.e. a warp's 32 lanes hold 32 consecutive channels of one pixel: coalesced NHWC stores
// without the shared-memory transpose.
//
// Questions this probe answers:
//   1. layout: does  tcgen05.cp.128x256b  of a no-swizzle K-major [128 x 16] fp16 tile (two 8-column core-matrix slabs,
//      LBO = slab pitch, SBO = 128 B -- the conv's operand image) followed by an A-from-TMEM MMA give the SAME accumulator
//      as the shared-memory MMA on the same tile?  (max |D_ts - D_ss| must be 0)
//   2. rate: issue interval of the A-from-TMEM MMA for N = 64 / 128 / 256 (expected 32 / 64 / 128 cycles: the math rate)
//   3. cost of tcgen05.cp per 4 KB tile (expected ~32 cycles of shared-memory read, overlappable with MMAs)
#include <cstdint>
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include "../../lidarcrafter_b200/csrc/common.cuh"

namespace b200 { void set_error(const char*, ...) {} bool pdl_enabled() { return false; } bool pdl_enabled_conv() { return false; } }
using namespace b200;

constexpr int SLAB = 130 * 16;   // bytes of one 8-channel group of a staged row (conv_tc.cu)

// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void tc_mma_f16_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc),
        "r"(acc)
        : "memory");
}
// 128 rows x 256 bits (= 16 fp16 of K) of a shared-memory matrix -> 128 lanes x 8 columns of tensor memory
__device__ __forceinline__ void tc_cp_128x256b(uint32_t d_tmem, uint64_t sdesc) {
    asm volatile("tcgen05.cp.cta_group::1.128x256b [%0], %1;" ::"r"(d_tmem), "l"(sdesc) : "memory");
}

__device__ __forceinline__ float val(int row, int k, int salt) {     // small integers: exact in fp16, sums exact in fp32
    return (float)(((row * 7 + k * 13 + salt * 5) % 9) - 4);
}

template <int N>
__global__ void __launch_bounds__(128, 1) ts_probe(int iters, float* maxdiff, unsigned long long* cyc) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    // A tile: weights [128 rows (channels)][16 k]; B tile: pixels [N rows][16 k]; both in the conv's operand image:
    // two 8-column slabs, row pitch 16 B inside a slab, slab pitch SLAB
    uint8_t* sa = smem;
    uint8_t* sb = smem + 8 * 1024;
    for (int i = threadIdx.x; i < 128 * 16; i += blockDim.x) {
        const int row = i / 16, k = i % 16;
        *reinterpret_cast<__half*>(sa + (k / 8) * SLAB + row * 16 + (k % 8) * 2) = __float2half(val(row, k, 1));
    }
    for (int i = threadIdx.x; i < N * 16; i += blockDim.x) {
        const int row = i / 16, k = i % 16;
        *reinterpret_cast<__half*>(sb + (k / 8) * (N * 16) + row * 16 + (k % 8) * 2) = __float2half(val(row, k, 2));
    }
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_slot), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_slot;
    const uint32_t d_ss = tm, d_ts = tm + (N <= 128 ? N : 0), a_tm = tm + 2 * 256 - 16;   // N = 256: one accumulator only
    const uint32_t idesc = make_idesc_f16(128, N);
    const uint64_t adesc = make_smem_desc(smem_u32(sa), SLAB, 128);
    const uint64_t bdesc = make_smem_desc(smem_u32(sb), N * 16, 128);
    uint32_t phase = 0;
    if (threadIdx.x == 0) {
        // ---- 1. layout check ----
        tc_mma_f16(d_ss, adesc, bdesc, idesc, 0);
        tc_cp_128x256b(a_tm, adesc);
        if (N <= 128) tc_mma_f16_ts(d_ts, a_tm, bdesc, idesc, 0);
        tc_commit(smem_u32(&bar));
    }
    mbar_wait(smem_u32(&bar), phase);
    phase ^= 1;
    tc_fence_after();
    if (N <= 128) {
        // thread t = TMEM lane t (warp w reads lanes 32w..32w+31): compare the two accumulators column by column
        float md = 0.f, ref_err = 0.f;
        for (int c0 = 0; c0 < N; c0 += 32) {
            float v0[32], v1[32];
            const uint32_t lane_base = (uint32_t)((threadIdx.x >> 5) * 32) << 16;
            tmem_ld_32x32(d_ss + lane_base + c0, v0);
            tmem_ld_32x32(d_ts + lane_base + c0, v1);
            for (int j = 0; j < 32; ++j) {
                md = fmaxf(md, fabsf(v0[j] - v1[j]));
                float r = 0.f;
                for (int k = 0; k < 16; ++k) r += val(threadIdx.x, k, 1) * val(c0 + j, k, 2);
                ref_err = fmaxf(ref_err, fabsf(v0[j] - r));
            }
        }
        atomicMax(reinterpret_cast<int*>(maxdiff), __float_as_int(md));          // non-negative floats order like ints
        atomicMax(reinterpret_cast<int*>(maxdiff) + 1, __float_as_int(ref_err));
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if (threadIdx.x == 0) {
        // ---- 2. issue interval of the A-from-TMEM MMA ----
        long long t0 = clock64();
        for (int it = 0; it < iters; ++it) tc_mma_f16_ts(d_ss, a_tm, bdesc, idesc, 1);
        tc_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), phase);
        phase ^= 1;
        cyc[0] = clock64() - t0;
        // same count of shared-memory-A MMAs for reference
        t0 = clock64();
        for (int it = 0; it < iters; ++it) tc_mma_f16(d_ss, adesc, bdesc, idesc, 1);
        tc_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), phase);
        phase ^= 1;
        cyc[1] = clock64() - t0;
        // ---- 3. tcgen05.cp alone, and interleaved 1 : 9 with MMAs (one weight tile per 9 taps) ----
        t0 = clock64();
        for (int it = 0; it < iters; ++it) tc_cp_128x256b(a_tm, adesc);
        tc_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), phase);
        phase ^= 1;
        cyc[2] = clock64() - t0;
        t0 = clock64();
        for (int it = 0; it < iters; ++it) {
            if (it % 9 == 0) tc_cp_128x256b(a_tm - 8 * ((it / 9) & 1), adesc);
            tc_mma_f16_ts(d_ss, a_tm, bdesc, idesc, 1);
        }
        tc_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), phase);
        cyc[3] = clock64() - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

template <int N>
static void run(int iters) {
    float* md;
    unsigned long long* cyc;
    cudaMalloc(&md, 8);
    cudaMalloc(&cyc, 32);
    cudaMemset(md, 0, 8);
    cudaFuncSetAttribute(ts_probe<N>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    ts_probe<N><<<1, 128, 64 * 1024>>>(iters, md, cyc);
    const cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("N=%d failed: %s\n", N, cudaGetErrorString(e)); return; }
    float h[2];
    unsigned long long c[4];
    cudaMemcpy(h, md, 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(c, cyc, 32, cudaMemcpyDeviceToHost);
    printf("N=%3d  max|D_ts - D_ss| = %g   max|D_ss - exact| = %g   cycles/MMA: A in TMEM %.1f, A in smem %.1f; tcgen05.cp %.1f; "
           "1 cp per 9 TS MMAs %.1f\n", N, h[0], h[1], (double)c[0] / iters, (double)c[1] / iters, (double)c[2] / iters,
           (double)c[3] / iters);
    cudaFree(md);
    cudaFree(cyc);
}

int main() {
    run<64>(4000);
    run<128>(4000);
    run<256>(2000);
    return 0;
}
