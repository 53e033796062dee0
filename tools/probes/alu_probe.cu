// Throughput probe for the elementwise ops of the fused conv front end (GroupNorm-apply + SiLU + fp16 hi/lo split):
// cycles per warp-instruction of MUFU.EX2 / MUFU.RCP / F2FP / HADD2.F32 / FFMA and of the whole per-element chain,
// with W warps per SM (one CTA per SM).   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o alu_probe alu_probe.cu
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int OP>
__global__ void probe(float* out, long long* cyc, int iters, float seed) {
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = seed + threadIdx.x * 1e-3f + i;
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 1) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(v[i]));
            if (OP == 2) {   // F2FP pack + unpack low
                unsigned h;
                asm volatile("cvt.rn.f16x2.f32 %0, %1, %1;" : "=r"(h) : "f"(v[i]));
                v[i] = __uint_as_float(h);
            }
            if (OP == 3) {   // HADD2.F32 (half -> float)
                __half hh = __ushort_as_half((unsigned short)__float_as_uint(v[i]));
                asm volatile("cvt.f32.f16 %0, %1;" : "=f"(v[i]) : "h"(__half_as_ushort(hh)));
            }
            if (OP == 4) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(v[i]));
            if (OP == 5) {   // the whole chain of one element: affine, silu, split
                float y = fmaf(v[i], 1.0001f, 0.001f);
                float e, r;
                asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(y * -1.4426950408889634f));
                asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.0f + e));
                y = y * r;
                const __half h = __float2half_rn(y);
                const float lo = y - __half2float(h);
                const __half l = __float2half_rn(lo);
                v[i] = __half2float(h) + __half2float(l);
            }
        }
    }
    const long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

int main() {
    float* out;
    long long* cyc;
    cudaMalloc(&out, 148 * 1024 * 4);
    cudaMalloc(&cyc, 148 * 8);
    const int iters = 2000;
    const char* names[] = {"MUFU.EX2", "MUFU.RCP", "F2FP.pack", "HADD2.F32", "FFMA", "silu+split chain (per element)"};
    for (int warps : {4, 8, 16, 32}) {
        for (int op = 0; op < 6; ++op) {
            switch (op) {
                case 0: probe<0><<<148, warps * 32>>>(out, cyc, iters, 0.5f); break;
                case 1: probe<1><<<148, warps * 32>>>(out, cyc, iters, 0.5f); break;
                case 2: probe<2><<<148, warps * 32>>>(out, cyc, iters, 0.5f); break;
                case 3: probe<3><<<148, warps * 32>>>(out, cyc, iters, 0.5f); break;
                case 4: probe<4><<<148, warps * 32>>>(out, cyc, iters, 0.5f); break;
                case 5: probe<5><<<148, warps * 32>>>(out, cyc, iters, 0.5f); break;
            }
            long long h[148];
            cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0;
            for (int i = 0; i < 148; ++i) avg += h[i];
            avg /= 148;
            const double per_sm_ops = (double)warps * iters * 8;   // warp-instructions (or elements x 32) per SM
            printf("warps/SM %2d  %-32s cycles %9.0f  -> %.2f cycles per warp-op per SMSP (%.1f lanes/clk/SM)\n", warps, names[op], avg,
                   avg / (per_sm_ops / 4), per_sm_ops * 32 / avg);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
