// K6 / K7 and LiDARUtility: the point-cloud side of the hot path.
//   * range_project      : dataset/transforms_3d/common.py:26-91 (load_points_as_images, scan_unfolding=False)
//   * points_in_boxes    : ops/roiaware_pool3d/src/roiaware_pool3d.cpp:121-168 (CPU semantics, MARGIN 1e-2)
//   * points_in_boxes_first / voxel_index : ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu:23-75,313-336
//   * depth_to_xyz       : utils/lidar.py:61-128 (denormalize -> revert_depth -> to_xyz)
// Integer outputs must be bit-exact against oracle/lidar_ops.c, so every fp32 operation that feeds a floor()/
// comparison is written with explicit round-to-nearest intrinsics (no FMA contraction) and the transcendental
// functions are evaluated in fp64 and rounded once to fp32 (the oracle defines them the same way).
#include "common.cuh"

namespace b200 {

// fp32 transcendental := round_to_float(fp64 function) -- identical definition in oracle/lidar_ops.c
__device__ __forceinline__ float asin_f(float x) { return (float)asin((double)x); }
__device__ __forceinline__ float atan2_f(float y, float x) { return (float)atan2((double)y, (double)x); }
__device__ __forceinline__ float cos_f(float x) { return (float)cos((double)x); }
__device__ __forceinline__ float sin_f(float x) { return (float)sin((double)x); }

struct ProjParams {
    const float* points;
    const int* npts;
    float* out;
    int* grid;
    unsigned long long* zbuf;
    int F, M, H, W;
    float min_depth, max_depth;
    double h_up, h_down;  // radians (np.deg2rad of python floats -> fp64)
};

__device__ __forceinline__ void project_point(const ProjParams& p, float x, float y, float z, float& depth, int& gh,
                                              int& gw) {
    // depth = ||xyz||_2 : np.linalg.norm on float32 -> ((x*x + y*y) + z*z) then sqrt, all fp32
    depth = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
    // elevation = arcsin(z / (depth + 1e-6)) [fp32]  + |h_down| [fp64]   (NumPy 2 promotion)
    const float ratio = __fdiv_rn(z, __fadd_rn(depth, 1e-6f));
    const double elev = (double)asin_f(ratio) + fabs(p.h_down);
    double g = 1.0 - elev / (p.h_up - p.h_down);
    g = floor(g * (double)p.H);
    g = fmin(fmax(g, 0.0), (double)(p.H - 1));
    gh = (int)g;
    // azimuth = -arctan2(y, x); grid_w = ((azimuth / pi + 1) / 2) % 1, all fp32
    const float az = -atan2_f(y, x);
    float t = __fdiv_rn(az, 3.14159274101257324f);
    t = __fadd_rn(t, 1.0f);
    t = __fmul_rn(t, 0.5f);
    t = fmodf(t, 1.0f);
    if (t < 0.f) t = __fadd_rn(t, 1.0f);
    float gwf = floorf(__fmul_rn(t, (float)p.W));
    gwf = fminf(fmaxf(gwf, 0.f), (float)(p.W - 1));
    gw = (int)gwf;
}

__global__ void proj_scatter_kernel(const ProjParams p) {
    const int f = blockIdx.y;
    const int n = p.npts ? p.npts[f] : p.M;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float4 pt = *reinterpret_cast<const float4*>(p.points + ((size_t)f * p.M + i) * 4);
        float depth;
        int gh, gw;
        project_point(p, pt.x, pt.y, pt.z, depth, gh, gw);
        if (p.grid) {
            p.grid[((size_t)f * p.M + i) * 2] = gh;
            p.grid[((size_t)f * p.M + i) * 2 + 1] = gw;
        }
        // nearest point wins; equal depth -> highest index wins (reference: last write of the argsort order)
        const unsigned long long key =
            ((unsigned long long)__float_as_uint(depth) << 32) | (unsigned long long)(0xFFFFFFFFu - (unsigned)i);
        atomicMin(p.zbuf + ((size_t)f * p.H + gh) * p.W + gw, key);
    }
}

__global__ void proj_gather_kernel(const ProjParams p) {
    const int f = blockIdx.y;
    const int hw = p.H * p.W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
        const unsigned long long key = p.zbuf[(size_t)f * hw + i];
        float* o = p.out + ((size_t)f * hw + i) * 6;
        if (key == 0xFFFFFFFFFFFFFFFFull) {
#pragma unroll
            for (int c = 0; c < 6; ++c) o[c] = 0.f;
        } else {
            const unsigned idx = 0xFFFFFFFFu - (unsigned)(key & 0xFFFFFFFFull);
            const float depth = __uint_as_float((unsigned)(key >> 32));
            const float4 pt = *reinterpret_cast<const float4*>(p.points + ((size_t)f * p.M + idx) * 4);
            o[0] = pt.x; o[1] = pt.y; o[2] = pt.z; o[3] = pt.w;
            o[4] = depth;
            o[5] = (depth >= p.min_depth && depth <= p.max_depth) ? 1.f : 0.f;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// points in boxes
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ int pt_in_box(float x, float y, float z, const float* bx, float margin, float& lx,
                                         float& ly) {
    const float cx = bx[0], cy = bx[1], cz = bx[2], dx = bx[3], dy = bx[4], dz = bx[5], rz = bx[6];
    if ((double)fabsf(__fsub_rn(z, cz)) > (double)dz / 2.0) return 0;
    const float sx = __fsub_rn(x, cx), sy = __fsub_rn(y, cy);
    const float cosa = cos_f(-rz), sina = sin_f(-rz);
    lx = __fadd_rn(__fmul_rn(sx, cosa), __fmul_rn(sy, -sina));
    ly = __fadd_rn(__fmul_rn(sx, sina), __fmul_rn(sy, cosa));
    const bool in = ((double)fabsf(lx) < (double)dx / 2.0 + (double)margin) &&
                    ((double)fabsf(ly) < (double)dy / 2.0 + (double)margin);
    return in ? 1 : 0;
}

__global__ void pib_all_kernel(const float* __restrict__ pts, const float* __restrict__ boxes, int* __restrict__ out,
                               int N, int M) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= M) return;
    float lx, ly;
    out[(size_t)i * M + j] = pt_in_box(pts[j * 3], pts[j * 3 + 1], pts[j * 3 + 2], boxes + i * 7, 1e-2f, lx, ly);
}

__global__ void pib_first_kernel(const float* __restrict__ pts, const float* __restrict__ boxes, int* __restrict__ out,
                                 int N, int M) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int b = blockIdx.y;
    if (j >= M) return;
    const float* pt = pts + ((size_t)b * M + j) * 3;
    float lx, ly;
    int r = -1;
    for (int k = 0; k < N; ++k) {
        if (pt_in_box(pt[0], pt[1], pt[2], boxes + ((size_t)b * N + k) * 7, 1e-5f, lx, ly)) {
            r = k;
            break;
        }
    }
    out[(size_t)b * M + j] = r;
}

__global__ void voxel_index_kernel(const float* __restrict__ pts, const float* __restrict__ rois, int* __restrict__ out,
                                   int N, int M, int ox, int oy, int oz) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= M) return;
    const float* bx = rois + i * 7;
    float lx = 0.f, ly = 0.f;
    int code = -1;
    if (pt_in_box(pts[j * 3], pts[j * 3 + 1], pts[j * 3 + 2], bx, 1e-5f, lx, ly)) {
        const float lz = __fsub_rn(pts[j * 3 + 2], bx[2]);
        const float dx = bx[3], dy = bx[4], dz = bx[5];
        const float xr = __fdiv_rn(dx, (float)ox), yr = __fdiv_rn(dy, (float)oy), zr = __fdiv_rn(dz, (float)oz);
        // unsigned idx = int(f) : truncation toward zero, then (as unsigned) min(max(.,0), n-1)
        unsigned xi = (unsigned)(int)__fdiv_rn(__fadd_rn(lx, __fdiv_rn(dx, 2.f)), xr);
        unsigned yi = (unsigned)(int)__fdiv_rn(__fadd_rn(ly, __fdiv_rn(dy, 2.f)), yr);
        unsigned zi = (unsigned)(int)__fdiv_rn(__fadd_rn(lz, __fdiv_rn(dz, 2.f)), zr);
        xi = min(max(xi, 0u), (unsigned)(ox - 1));
        yi = min(max(yi, 0u), (unsigned)(oy - 1));
        zi = min(max(zi, 0u), (unsigned)(oz - 1));
        code = (int)((xi << 16) + (yi << 8) + zi);
    }
    out[(size_t)i * M + j] = code;
}

// ---------------------------------------------------------------------------------------------------------
// LiDARUtility: x in [-1,1] -> denormalize -> revert_depth (log_depth) -> mask -> xyz
// ---------------------------------------------------------------------------------------------------------
__global__ void depth_to_xyz_kernel(const float* __restrict__ xn, const float* __restrict__ ang,
                                    float* __restrict__ depth, float* __restrict__ xyz, int HW, float min_d,
                                    float max_d, float log2_max) {
    const int b = blockIdx.y;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
        const float nd = (xn[(size_t)b * HW + i] + 1.f) * 0.5f;
        float m = exp2f(nd * log2_max) - 1.f;
        const float mask = (m > min_d && m < max_d) ? 1.f : 0.f;
        m *= mask;
        if (depth) depth[(size_t)b * HW + i] = m;
        if (xyz) {
            const float phi = ang[i], th = ang[HW + i];
            const float m2 = (m > min_d && m < max_d) ? 1.f : 0.f;
            const float cp = cosf(phi);
            xyz[((size_t)b * 3 + 0) * HW + i] = m * cp * cosf(th) * m2;
            xyz[((size_t)b * 3 + 1) * HW + i] = m * cp * sinf(th) * m2;
            xyz[((size_t)b * 3 + 2) * HW + i] = m * sinf(phi) * m2;
        }
    }
}

}  // namespace b200

using namespace b200;

extern "C" int b200_range_project(const float* points, const int* npts, float* out, int* grid, void* zbuf, int F, int M,
                                  int H, int W, float min_depth, float max_depth, float fov_up_deg, float fov_down_deg,
                                  void* stream) {
    B200_CHECK_ARG(points && out && zbuf && F > 0 && M > 0 && H > 0 && W > 0);
    cudaStream_t st = (cudaStream_t)stream;
    const double d2r = 3.14159265358979323846 / 180.0;
    ProjParams p{points, npts, out, grid, (unsigned long long*)zbuf, F, M, H, W, min_depth, max_depth,
                 (double)fov_up_deg * d2r, (double)fov_down_deg * d2r};
    if (cudaMemsetAsync(zbuf, 0xFF, (size_t)F * H * W * 8, st) != cudaSuccess) {
        set_error("range_project: memset failed");
        return B200_E_CUDA;
    }
    dim3 g1(cdiv(M, 256), F);
    proj_scatter_kernel<<<g1, 256, 0, st>>>(p);
    B200_CHECK_LAUNCH();
    dim3 g2(cdiv(H * W, 256), F);
    proj_gather_kernel<<<g2, 256, 0, st>>>(p);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

extern "C" int b200_points_in_boxes(const float* pts, const float* boxes, int* out, int N, int M, void* stream) {
    B200_CHECK_ARG(pts && boxes && out && N > 0 && M > 0 && N <= 65535);
    dim3 g(cdiv(M, 256), N);
    pib_all_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(pts, boxes, out, N, M);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

extern "C" int b200_points_in_boxes_first(const float* pts, const float* boxes, int* out, int B, int N, int M,
                                          void* stream) {
    B200_CHECK_ARG(pts && boxes && out && B > 0 && N >= 0 && M > 0 && B <= 65535);
    dim3 g(cdiv(M, 256), B);
    pib_first_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(pts, boxes, out, N, M);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

extern "C" int b200_voxel_index(const float* pts, const float* rois, int* out, int N, int M, int out_x, int out_y,
                                int out_z, void* stream) {
    B200_CHECK_ARG(pts && rois && out && N > 0 && M > 0 && N <= 65535);
    B200_CHECK_ARG(out_x > 0 && out_y > 0 && out_z > 0 && out_x <= 256 && out_y <= 256 && out_z <= 256);
    dim3 g(cdiv(M, 256), N);
    voxel_index_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(pts, rois, out, N, M, out_x, out_y, out_z);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

extern "C" int b200_depth_to_xyz(const float* x_norm, const float* ray_angles, float* depth, float* xyz, int B, int H,
                                 int W, float min_depth, float max_depth, void* stream) {
    B200_CHECK_ARG(x_norm && (depth || xyz) && (!xyz || ray_angles));
    dim3 g(cdiv(H * W, 256), B);
    depth_to_xyz_kernel<<<g, 256, 0, (cudaStream_t)stream>>>(x_norm, ray_angles, depth, xyz, H * W, min_depth, max_depth,
                                                            log2f(max_depth + 1.f));
    B200_CHECK_LAUNCH();
    return B200_OK;
}
