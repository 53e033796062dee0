// Microbenchmark (run on the GPU box): issue rate of tcgen05.mma as a function of the shared-memory operand geometry.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_probe mma_probe.cu && ./mma_probe
// Question it answers: does an A operand whose 128-byte core matrices are NOT 128-byte aligned (the dx*16-byte tap
// shift / the 2080-byte slab pitch of conv_tc.cu) cost extra shared-memory wavefronts per MMA?
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../lidarcrafter_b200/csrc/common.cuh"

namespace b200 { void set_error(const char*, ...) {} bool pdl_enabled() { return false; } }
using namespace b200;

struct Cfg {
    int N;            // MMA N
    int kind;         // 0 = f16 (K=16), 1 = f8f6f4 (K=32)
    uint32_t a_off;   // start-address offset of A (bytes)
    uint32_t a_lbo;   // K-direction core-matrix stride of A (bytes)
    int swz;          // 0 = no swizzle (core-matrix layout), 1 = SWIZZLE_128B K-major (rows of 128 B, SBO 1024)
    int n_a;          // distinct A tiles cycled through (different rows/taps)
    uint32_t a_step;  // byte step between the distinct A tiles
    int n_acc;        // distinct TMEM accumulators cycled through (dependent-accumulate chains in flight)
};

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)1 << 16;                       // LBO (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;             // SBO: 8 rows x 128 B
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)((saddr >> 7) & 7) << 49;      // base offset for starts that are not 1024-B aligned
    d |= (uint64_t)2 << 61;                       // SWIZZLE_128B
    return d;
}

template <int KIND, int NACC>
__global__ void __launch_bounds__(128, 1) probe(Cfg c, int iters, unsigned long long* out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar;
    for (int i = threadIdx.x; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u;
    if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); fence_barrier_init(); }
    if (threadIdx.x < 32) tmem_alloc(smem_u32(&tmem_slot), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_slot;
    if (threadIdx.x == 0) {
        const uint32_t sa = smem_u32(smem), sb = sa + 160 * 1024;
        const uint32_t idesc = make_idesc_f16(128, c.N);
        const uint64_t bdesc = c.swz ? desc_sw128(sb) : make_smem_desc(sb, c.N * 16, 128);
        uint64_t ad[8];
        for (int j = 0; j < 8; ++j) {
            const uint32_t a = sa + c.a_off + (j % c.n_a) * c.a_step;
            ad[j] = c.swz ? desc_sw128(a) : make_smem_desc(a, c.a_lbo, 128);
        }
        uint32_t accv[8];
        for (int j = 0; j < 8; ++j) accv[j] = tm + (j % NACC) * c.N;
        const long long t0 = clock64();
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                if (KIND == 0)
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(accv[j]), "l"(ad[j]), "l"(bdesc), "r"(idesc));
                else
                    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, 1, 0;\n\ttcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}"
                                 ::"r"(accv[j]), "l"(ad[j]), "l"(bdesc), "r"(idesc));
            }
        }
        const long long t1 = clock64();
        tc_commit(smem_u32(&bar));
        mbar_wait(smem_u32(&bar), 0);
        out[blockIdx.x] = clock64() - t0;
        out[gridDim.x + blockIdx.x] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (threadIdx.x < 32) tmem_dealloc(tm, 512);
}

int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    #define S(K_, A_) cudaFuncSetAttribute(probe<K_, A_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    S(0, 1) S(0, 2) S(0, 4) S(0, 8) S(1, 1) S(1, 2) S(1, 4) S(1, 8)
#undef S
    unsigned long long* d;
    cudaMalloc(&d, 2 * sms * 8);
    const Cfg cfgs[] = {
        // N kind a_off a_lbo swz n_a a_step n_acc
        {64, 0, 0, 2048, 0, 1, 0, 1},   {64, 0, 0, 2048, 0, 1, 0, 2},   {64, 0, 0, 2048, 0, 1, 0, 4},   {64, 0, 0, 2048, 0, 1, 0, 8},
        {128, 0, 0, 2048, 0, 1, 0, 1},  {128, 0, 0, 2048, 0, 1, 0, 2},  {128, 0, 0, 2048, 0, 1, 0, 4},
        {256, 0, 0, 2048, 0, 1, 0, 1},  {256, 0, 0, 2048, 0, 1, 0, 2},
        // alignment of the A core matrices (8 accumulators / 4 accumulators in flight)
        {64, 0, 16, 2048, 0, 1, 0, 8},  {64, 0, 32, 2048, 0, 1, 0, 8},  {64, 0, 64, 2048, 0, 1, 0, 8},
        {64, 0, 0, 2080, 0, 1, 0, 8},   {64, 0, 16, 2080, 0, 1, 0, 8},  {64, 0, 0, 2176, 0, 1, 0, 8},
        {128, 0, 16, 2048, 0, 1, 0, 4}, {128, 0, 0, 2080, 0, 1, 0, 4},  {128, 0, 16, 2080, 0, 1, 0, 4}, {256, 0, 16, 2080, 0, 1, 0, 2},
        // distinct A tiles
        {64, 0, 0, 2080, 0, 8, 4160, 8}, {64, 0, 16, 2080, 0, 8, 4160, 8}, {128, 0, 16, 2080, 0, 8, 4160, 4}, {64, 0, 0, 2048, 0, 8, 4096, 8},
        // fp8 kind (K = 32)
        {64, 1, 0, 2048, 0, 1, 0, 8},   {128, 1, 0, 2048, 0, 1, 0, 4},  {64, 1, 16, 2080, 0, 1, 0, 8},  {128, 1, 16, 2080, 0, 1, 0, 4}, {256, 1, 0, 2048, 0, 1, 0, 2},
        // SWIZZLE_128B K-major rows (128 B per pixel): aligned start, +128 B (one pixel), +32 B (next K step)
        {64, 0, 0, 0, 1, 1, 0, 8},      {128, 0, 0, 0, 1, 1, 0, 4},     {256, 0, 0, 0, 1, 1, 0, 2},
        {64, 0, 128, 0, 1, 1, 0, 8},    {128, 0, 128, 0, 1, 1, 0, 4},   {64, 0, 32, 0, 1, 1, 0, 8},     {128, 0, 160, 0, 1, 1, 0, 4},
        {64, 1, 0, 0, 1, 1, 0, 8},      {128, 1, 128, 0, 1, 1, 0, 4},
    };
    const int iters = 2000;
    printf("%5s %4s %6s %6s %3s %3s %4s | cycles/MMA  (TFLOP/s/GPU at 1.9 GHz)\n", "N", "kind", "a_off", "a_lbo", "swz", "n_a", "nacc");
    for (const Cfg& c : cfgs) {
        auto launch = [&]() {
#define L(K_, A_) if (c.kind == K_ && c.n_acc == A_) probe<K_, A_><<<sms, 128, 200 * 1024>>>(c, iters, d);
            L(0, 1) L(0, 2) L(0, 4) L(0, 8) L(1, 1) L(1, 2) L(1, 4) L(1, 8)
#undef L
        };
        launch();
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("cfg failed: %s\n", cudaGetErrorString(e)); return 1; }
        launch();
        cudaDeviceSynchronize();
        unsigned long long h[512];
        cudaMemcpy(h, d, 2 * sms * 8, cudaMemcpyDeviceToHost);
        double s = 0, si = 0;
        for (int i = 0; i < sms; ++i) { s += (double)h[i]; si += (double)h[sms + i]; }
        const double cyc = s / sms / ((double)iters * 8);
        const double cyc_issue = si / sms / ((double)iters * 8);
        const double flop = 2.0 * 128 * c.N * (c.kind ? 32 : 16);
        printf("%5d %4s %6u %6u %3d %3d %4d | %8.1f  issue %6.1f  (%7.0f)\n", c.N, c.kind ? "f8" : "f16", c.a_off, c.a_lbo, c.swz,
               c.n_a, c.n_acc, cyc, cyc_issue, flop / cyc * 1.9e9 * sms / 1e12);
    }
    return 0;
}
