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

// points_in_boxes_gpu / generate_pts_mask_for_box3d exist ONLY as CUDA in the reference (roiaware_pool3d_kernel.cu:15-36): its
// build is nvcc with default flags, i.e. the products of the local-frame rotation are contracted into FMAs and cos / sin of
// a float resolve to the CUDA cosf / sinf.  The two kernels that replace them follow THAT arithmetic -- the expression
// below is left to the compiler exactly like the reference's (same toolkit -> same contraction, same libdevice) -- and are
// pinned bit for bit against oracle/_ref/libref_roiaware_cuda.so on the GPU box; pt_in_box() above (explicit
// round-to-nearest products, cos / sin through fp64) is the arithmetic of the reference's C++ points_in_boxes_cpu.
__device__ __forceinline__ int pt_in_box_fma(const float* pt, const float* bx, float& lx, float& ly) {
    const float margin = 1e-5;
    const float px = pt[0], py = pt[1], pz = pt[2];
    const float cx = bx[0], cy = bx[1], cz = bx[2], dx = bx[3], dy = bx[4], dz = bx[5], rz = bx[6];
    if (fabsf(pz - cz) > dz / 2.0) return 0;
    const float sx = px - cx, sy = py - cy;
    const float ca = cosf(-rz), sa = sinf(-rz);
    lx = sx * ca + sy * (-sa);
    ly = sx * sa + sy * ca;
    const float in = (fabs(lx) < dx / 2.0 + margin) & (fabs(ly) < dy / 2.0 + margin);
    return in;
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
        if (pt_in_box_fma(pt, boxes + ((size_t)b * N + k) * 7, lx, ly)) {
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
    if (pt_in_box_fma(pts + j * 3, bx, lx, ly) > 0) {
        const float lz = pts[j * 3 + 2] - bx[2];
        const float dx = bx[3], dy = bx[4], dz = bx[5];
        const float xr = dx / ox, yr = dy / oy, zr = dz / oz;
        // unsigned idx = int(f) : truncation toward zero, then (as unsigned) min(max(., 0), n - 1)
        unsigned xi = int((lx + dx / 2) / xr);
        unsigned yi = int((ly + dy / 2) / yr);
        unsigned zi = int((lz + dz / 2) / zr);
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

// ---------------------------------------------------------------------------------------------------------
// range projection of float64 points (the temporal glue re-projects the ego-motion-warped background, a float64 array:
// tools/vis_tools/utils/pipe_related.py:244-255 -> CustomDataset -> load_points_as_images; every intermediate is fp64
// there and only the final image is cast to fp32).  The 64-bit depth leaves no room for the point index in one atomic
// key, so: pass 1 atomicMin(depth bits), pass 2 atomicMax(index) among the points AT the minimum, pass 3 gather.
// ---------------------------------------------------------------------------------------------------------
struct ProjParams64 {
    const double* points;
    const int* npts;
    float* out;
    unsigned long long* zbuf;
    int* win;
    int F, M, H, W;
    double min_depth, max_depth;
    double h_up, h_down;
};

__device__ __forceinline__ void project_point64(const ProjParams64& p, double x, double y, double z, double& depth, int& gh,
                                                int& gw) {
    depth = __dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z)));
    const double elev = asin(__ddiv_rn(z, __dadd_rn(depth, 1e-6))) + fabs(p.h_down);
    double g = __dsub_rn(1.0, __ddiv_rn(elev, __dsub_rn(p.h_up, p.h_down)));
    g = floor(__dmul_rn(g, (double)p.H));
    gh = (int)fmin(fmax(g, 0.0), (double)(p.H - 1));
    const double az = -atan2(y, x);
    double t = __dmul_rn(__dadd_rn(__ddiv_rn(az, 3.141592653589793), 1.0), 0.5);      // (az / pi + 1) / 2
    t = fmod(t, 1.0);
    if (t < 0.0) t = __dadd_rn(t, 1.0);
    const double gwf = floor(__dmul_rn(t, (double)p.W));
    gw = (int)fmin(fmax(gwf, 0.0), (double)(p.W - 1));
}

template <int PASS>
__global__ void proj64_scatter_kernel(const ProjParams64 p) {
    const int f = blockIdx.y;
    const int n = p.npts ? p.npts[f] : p.M;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const double* pt = p.points + ((size_t)f * p.M + i) * 4;
        double depth;
        int gh, gw;
        project_point64(p, pt[0], pt[1], pt[2], depth, gh, gw);
        const size_t px = ((size_t)f * p.H + gh) * p.W + gw;
        const unsigned long long key = (unsigned long long)__double_as_longlong(depth);   // depth >= 0: bits order like values
        if (PASS == 0) atomicMin(p.zbuf + px, key);
        else if (p.zbuf[px] == key) atomicMax(p.win + px, i);     // equal depth: highest index wins (as in the fp32 kernel)
    }
}

__global__ void proj64_gather_kernel(const ProjParams64 p) {
    const int f = blockIdx.y;
    const int hw = p.H * p.W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
        const int idx = p.win[(size_t)f * hw + i];
        float* o = p.out + ((size_t)f * hw + i) * 6;
        if (idx < 0) {
#pragma unroll
            for (int c = 0; c < 6; ++c) o[c] = 0.f;
        } else {
            const double* pt = p.points + ((size_t)f * p.M + idx) * 4;
            const double depth = __longlong_as_double((long long)p.zbuf[(size_t)f * hw + i]);
            o[0] = (float)pt[0]; o[1] = (float)pt[1]; o[2] = (float)pt[2]; o[3] = (float)pt[3];
            o[4] = (float)depth;
            o[5] = (depth >= p.min_depth && depth <= p.max_depth) ? 1.f : 0.f;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------
// 3-D boxes -> 2-D boxes / condition mask / loss-weight map (dataset/transforms_3d/common.py:99-216, convert_boxes_to_2d)
//   boxes [F,N,8] (x, y, z, l, w, h, yaw, class) as fp32 or fp64 -- the reference's dtype flow depends on it: corner
//   offsets, centre, centre depth, cos / sin of the yaw are evaluated in the boxes' dtype, the rotation and the projection
//   of the 8 corners in fp64.  Pixel rectangles are int(x * W) of the fp64 grid coordinates; later boxes overwrite earlier
//   ones; a rectangle wider than 0.6 W straddles the azimuth seam and is drawn as [0, x1) + [x2, W).
// ---------------------------------------------------------------------------------------------------------
struct BoxRec { int x1, y1, x2, y2, wrap; float cls, depth, weight; };

template <typename T>
__device__ __forceinline__ T box_cos(T a);
template <> __device__ __forceinline__ float box_cos<float>(float a) { return cos_f(a); }
template <> __device__ __forceinline__ double box_cos<double>(double a) { return cos(a); }
template <typename T>
__device__ __forceinline__ T box_sin(T a);
template <> __device__ __forceinline__ float box_sin<float>(float a) { return sin_f(a); }
template <> __device__ __forceinline__ double box_sin<double>(double a) { return sin(a); }

__device__ __forceinline__ float mul_rn(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ double mul_rn(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ float add_rn(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ double add_rn(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ float sqrt_rn(float a) { return __fsqrt_rn(a); }
__device__ __forceinline__ double sqrt_rn(double a) { return __dsqrt_rn(a); }

template <typename T>
__global__ void box_rect_kernel(const T* __restrict__ boxes, int N, int H, int W, double h_up, double h_down,
                                double* __restrict__ boxes_2d, BoxRec* __restrict__ rec) {
    extern __shared__ double s_uv[];          // [N * 8][2] normalised (grid_w / W, grid_h / H) of the corners
    __shared__ int s_max_area;
    const int f = blockIdx.x;
    const T* bf = boxes + (size_t)f * N * 8;
    if (threadIdx.x == 0) s_max_area = INT_MIN;
    for (int t = threadIdx.x; t < N * 8; t += blockDim.x) {
        const int i = t >> 3, k = t & 7;
        const T* b = bf + i * 8;
        // corner offsets in the boxes' dtype: x: +l/2 for corners 0,1,4,5; y: +w/2 for 0,3,4,7; z: +h/2 for 0..3
        const T hx = b[3] / (T)2, hy = b[4] / (T)2, hz = b[5] / (T)2;
        const double lx = (double)((k == 0 || k == 1 || k == 4 || k == 5) ? hx : -hx);
        const double ly = (double)((k == 0 || k == 3 || k == 4 || k == 7) ? hy : -hy);
        const double lz = (double)(k < 4 ? hz : -hz);
        const double c = (double)box_cos<T>(b[6]), s = (double)box_sin<T>(b[6]);
        // rotz @ corner (fp64; the zero entries of the rotation add exact zeros), + centre
        const double x = __dadd_rn(__dadd_rn(__dmul_rn(c, lx), __dmul_rn(-s, ly)), (double)b[0]);
        const double y = __dadd_rn(__dadd_rn(__dmul_rn(s, lx), __dmul_rn(c, ly)), (double)b[1]);
        const double z = __dadd_rn(lz, (double)b[2]);
        const double depth = __dadd_rn(__dsqrt_rn(__dadd_rn(__dadd_rn(__dmul_rn(x, x), __dmul_rn(y, y)), __dmul_rn(z, z))), 1e-6);
        const double elev = asin(__ddiv_rn(z, depth)) + fabs(h_down);
        double g = __dsub_rn(1.0, __ddiv_rn(elev, __dsub_rn(h_up, h_down)));
        g = fmin(fmax(floor(__dmul_rn(g, (double)H)), 0.0), (double)(H - 1));
        const double az = -atan2(y, x);
        double tw = __dmul_rn(__dadd_rn(__ddiv_rn(az, 3.141592653589793), 1.0), 0.5);
        tw = fmod(tw, 1.0);
        if (tw < 0.0) tw = __dadd_rn(tw, 1.0);
        const double gw = fmin(fmax(floor(__dmul_rn(tw, (double)W)), 0.0), (double)(W - 1));
        s_uv[2 * t] = __ddiv_rn(gw, (double)W);
        s_uv[2 * t + 1] = __ddiv_rn(g, (double)H);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        double u0 = s_uv[16 * i], v0 = s_uv[16 * i + 1], u1 = u0, v1 = v0;
        for (int k = 1; k < 8; ++k) {
            u0 = fmin(u0, s_uv[16 * i + 2 * k]); u1 = fmax(u1, s_uv[16 * i + 2 * k]);
            v0 = fmin(v0, s_uv[16 * i + 2 * k + 1]); v1 = fmax(v1, s_uv[16 * i + 2 * k + 1]);
        }
        double* o = boxes_2d + ((size_t)f * N + i) * 4;
        o[0] = u0; o[1] = v0; o[2] = u1; o[3] = v1;
        BoxRec r;
        r.x1 = (int)__dmul_rn(u0, (double)W); r.x2 = (int)__dmul_rn(u1, (double)W);
        r.y1 = (int)__dmul_rn(v0, (double)H); r.y2 = (int)__dmul_rn(v1, (double)H);
        r.wrap = ((double)(r.x2 - r.x1) / (double)W > 0.6) ? 1 : 0;
        const T* b = bf + i * 8;
        r.cls = (float)b[7];
        // centre depth in the boxes' dtype: ||(x, y, z)|| + 1e-6
        const T cd = sqrt_rn(add_rn(add_rn(mul_rn(b[0], b[0]), mul_rn(b[1], b[1])), mul_rn(b[2], b[2])));
        r.depth = (float)add_rn(cd, (T)1e-6);
        const int area = (r.wrap ? (W - r.x2 + r.x1) : (r.x2 - r.x1)) * (r.y2 - r.y1);
        r.weight = (float)area;          // finished below, once the largest area is known
        atomicMax(&s_max_area, area);
        rec[(size_t)f * N + i] = r;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        BoxRec& r = rec[(size_t)f * N + i];
        r.weight = __fsub_rn(3.f, __fdiv_rn(r.weight, (float)s_max_area));      // fp32: 3 - area / max(area)
    }
}

__global__ void box_raster_kernel(const BoxRec* __restrict__ rec, int N, int H, int W, float* __restrict__ mask,
                                  float* __restrict__ weight) {
    extern __shared__ BoxRec s_rec[];
    const int f = blockIdx.y;
    for (int i = threadIdx.x; i < N; i += blockDim.x) s_rec[i] = rec[(size_t)f * N + i];
    __syncthreads();
    const int hw = H * W;
    for (int px = blockIdx.x * blockDim.x + threadIdx.x; px < hw; px += gridDim.x * blockDim.x) {
        const int y = px / W, x = px - y * W;
        float cls = 0.f, dep = 0.f, wsum = 0.f;
        for (int i = 0; i < N; ++i) {
            const BoxRec& r = s_rec[i];
            const bool in = y >= r.y1 && y < r.y2 && (r.wrap ? (x < r.x1 || x >= r.x2) : (x >= r.x1 && x < r.x2));
            if (in) { cls = r.cls; dep = r.depth; }
            wsum = __fadd_rn(wsum, in ? r.weight : 0.f);
        }
        mask[((size_t)f * 2 + 0) * hw + px] = cls;
        mask[((size_t)f * 2 + 1) * hw + px] = dep;
        if (weight) weight[(size_t)f * hw + px] = expf(wsum);
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

extern "C" int b200_range_project_f64(const double* points, const int* npts, float* out, void* zbuf, int* winner, int F,
                                      int M, int H, int W, float min_depth, float max_depth, float fov_up_deg,
                                      float fov_down_deg, void* stream) {
    B200_CHECK_ARG(points && out && zbuf && winner && F > 0 && M > 0 && H > 0 && W > 0);
    cudaStream_t st = (cudaStream_t)stream;
    const double d2r = 3.14159265358979323846 / 180.0;
    ProjParams64 p{points, npts, out, (unsigned long long*)zbuf, winner, F, M, H, W, (double)min_depth, (double)max_depth,
                   (double)fov_up_deg * d2r, (double)fov_down_deg * d2r};
    if (cudaMemsetAsync(zbuf, 0xFF, (size_t)F * H * W * 8, st) != cudaSuccess ||
        cudaMemsetAsync(winner, 0xFF, (size_t)F * H * W * 4, st) != cudaSuccess) {
        set_error("range_project_f64: memset failed");
        return B200_E_CUDA;
    }
    dim3 g1(cdiv(M, 256), F);
    proj64_scatter_kernel<0><<<g1, 256, 0, st>>>(p);
    B200_CHECK_LAUNCH();
    proj64_scatter_kernel<1><<<g1, 256, 0, st>>>(p);
    B200_CHECK_LAUNCH();
    dim3 g2(cdiv(H * W, 256), F);
    proj64_gather_kernel<<<g2, 256, 0, st>>>(p);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

extern "C" size_t b200_boxes_to_mask_workspace(int F, int N) { return (size_t)F * N * sizeof(BoxRec); }

extern "C" int b200_boxes_to_mask(const void* boxes, int boxes_f64, int F, int N, int H, int W, float fov_up_deg,
                                  float fov_down_deg, double* boxes_2d, float* mask, float* weight, void* workspace,
                                  void* stream) {
    B200_CHECK_ARG(boxes && boxes_2d && mask && workspace && F > 0 && N > 0 && N <= 512 && H > 0 && W > 0 && F <= 65535);
    cudaStream_t st = (cudaStream_t)stream;
    const double d2r = 3.14159265358979323846 / 180.0;
    const double h_up = (double)fov_up_deg * d2r, h_down = (double)fov_down_deg * d2r;
    BoxRec* rec = (BoxRec*)workspace;
    if (boxes_f64)
        box_rect_kernel<double><<<F, 128, (size_t)N * 16 * sizeof(double), st>>>((const double*)boxes, N, H, W, h_up, h_down, boxes_2d, rec);
    else
        box_rect_kernel<float><<<F, 128, (size_t)N * 16 * sizeof(double), st>>>((const float*)boxes, N, H, W, h_up, h_down, boxes_2d, rec);
    B200_CHECK_LAUNCH();
    dim3 g(cdiv(H * W, 256), F);
    box_raster_kernel<<<g, 256, (size_t)N * sizeof(BoxRec), st>>>(rec, N, H, W, mask, weight);
    B200_CHECK_LAUNCH();
    return B200_OK;
}
