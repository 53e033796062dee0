// K8: the evaluation-side projection and the integer voxel quantiser.
//   * pcd2range        : lidargen/metrics/metric_utils.py:65-121   (strict depth test, -1 fill, no +1e-6 / no modulo)
//   * range2xyz        : lidargen/metrics/metric_utils.py:124-154  (fp64 ray directions)
//   * quantize_coords  : np.floor(coords / voxel_size).astype(np.int32) of sparse_quantize (:44-62) and of the
//                        pcd2bev_* / pcd2voxel_full callers (:170-306)
//   * ravel_hash       : metric_utils.py:28-41
//   * sparse_quantize  : np.unique(ravel_hash(coords), return_index, return_inverse) (:53-62)
//   * bev_occupancy_sum / voxel_occupancy : pcd2bev_sum (:231-256) / pcd2voxel_full (:170-199)
//
// All of it is HBM-/atomic-bound integer work.  np.unique is a sort in the reference; here the ravel hash IS a dense
// index into the clouds' bounding grid, so uniqueness and the sorted order come from a BITMAP of the grid: set bits,
// popcount-scan the words, rank(key) = set bits below key.  No comparison sort, one pass over the points for the bits,
// one for the ranks, streaming passes over the bitmap (32 MB for a 60 m x 60 m x 9 m nuScenes grid at 5 cm).
//
// Integer outputs must be bit-exact against oracle/metrics_ops.c: fp32 operations that feed floor()/comparisons use
// explicit round-to-nearest intrinsics (no FMA contraction); fp32 asin/atan2 := (float) f((double) x) like lidar_ops.cu.
#include "common.cuh"

namespace b200 {

__device__ __forceinline__ float m_asin_f(float x) { return (float)asin((double)x); }
__device__ __forceinline__ float m_atan2_f(float y, float x) { return (float)atan2((double)y, (double)x); }

// ---------------------------------------------------------------------------------------------------------
// pcd2range
// ---------------------------------------------------------------------------------------------------------
struct P2RParams {
    const float* pcd;        // [F, M, 3]
    const int* npts;         // [F] or null
    const float* feature;    // [F, M] or null
    float* proj_range;       // [F, H, W]
    float* proj_feature;     // [F, H, W] or null
    unsigned long long* zbuf;
    int F, M, H, W;
    float fov_down_abs, fov_range;   // fp32(|fov_down|), fp32(|fov_down| + |fov_up|)   (python floats are "weak": cast to fp32)
    float dmin, dmax, feature_fill;
};

__global__ void p2r_scatter_kernel(const P2RParams p) {
    const int f = blockIdx.y;
    const int n = p.npts ? p.npts[f] : p.M;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        const float* pt = p.pcd + ((size_t)f * p.M + i) * 3;
        const float x = pt[0], y = pt[1], z = pt[2];
        const float depth = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(x, x), __fmul_rn(y, y)), __fmul_rn(z, z)));
        if (!(depth > p.dmin && depth < p.dmax)) continue;
        const float yaw = -m_atan2_f(y, x);
        const float pitch = m_asin_f(__fdiv_rn(z, depth));
        float px = __fmul_rn(0.5f, __fadd_rn(__fdiv_rn(yaw, 3.14159274101257324f), 1.0f));
        float py = __fsub_rn(1.0f, __fdiv_rn(__fadd_rn(pitch, p.fov_down_abs), p.fov_range));
        px = __fmul_rn(px, (float)p.W);
        py = __fmul_rn(py, (float)p.H);
        const int ix = (int)fmaxf(0.f, fminf((float)(p.W - 1), floorf(px)));
        const int iy = (int)fmaxf(0.f, fminf((float)(p.H - 1), floorf(py)));
        // nearest wins; equal depth: lowest index wins (a stable descending-depth order writes it last)
        const unsigned long long key = ((unsigned long long)__float_as_uint(depth) << 32) | (unsigned)i;
        atomicMin(p.zbuf + ((size_t)f * p.H + iy) * p.W + ix, key);
    }
}

__global__ void p2r_gather_kernel(const P2RParams p) {
    const int f = blockIdx.y;
    const int hw = p.H * p.W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
        const unsigned long long key = p.zbuf[(size_t)f * hw + i];
        const bool hit = key != 0xFFFFFFFFFFFFFFFFull;
        p.proj_range[(size_t)f * hw + i] = hit ? __uint_as_float((unsigned)(key >> 32)) : -1.f;
        if (p.proj_feature)
            p.proj_feature[(size_t)f * hw + i] =
                hit ? p.feature[(size_t)f * p.M + (unsigned)(key & 0xFFFFFFFFull)] : p.feature_fill;
    }
}

// ---------------------------------------------------------------------------------------------------------
// range2xyz (fp64 like the reference: np.meshgrid -> float64 angles; depth stays the fp32 image value)
// ---------------------------------------------------------------------------------------------------------
__global__ void range2xyz_kernel(const float* __restrict__ img, double* __restrict__ xyz, int H, int W, double fov_down_abs,
                                 double fov_range, float dmin, float dmax, float depth_scale, int log_scale) {
    const int f = blockIdx.y;
    const int hw = H * W;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
        const int r = i / W, c = i - r * W;
        const float v = img[(size_t)f * hw + i];
        const float depth = log_scale ? __fsub_rn(exp2f(__fmul_rn(v, depth_scale)), 1.f) : v;
        const double sx = (double)c / (double)W, sy = (double)r / (double)H;
        const double yaw = 3.14159265358979323846 * (sx * 2.0 - 1.0);
        const double pitch = (1.0 - sy) * fov_range - fov_down_abs;
        const bool ok = depth > dmin && depth < dmax;
        double* o = xyz + (size_t)f * 3 * hw + i;
        const double cp = cos(pitch);
        o[0] = ok ? cos(yaw) * cp * (double)depth : -1.0;
        o[hw] = ok ? -sin(yaw) * cp * (double)depth : -1.0;
        o[2 * hw] = ok ? sin(pitch) * (double)depth : -1.0;
    }
}

// ---------------------------------------------------------------------------------------------------------
// voxel quantiser
// ---------------------------------------------------------------------------------------------------------
__global__ void minmax_init_kernel(int* mm) {
    if (threadIdx.x < 3) mm[threadIdx.x] = 0x7FFFFFFF;
    else if (threadIdx.x < 6) mm[threadIdx.x] = (int)0x80000000;
}

// voxel[i,d] = (int32) floor(coords[i,d] / vs[d]);  div_f32 = 1: the division is the fp32 one of `pcd / voxel_size`
// with a python-float voxel size (pcd2bev_*, pcd2voxel_full); 0: the fp64 one of `coords / np.array(voxel_size)`
// (sparse_quantize).  minmax[0:3] / [3:6] = per-axis min / max (block reduction + one atomic pair per block and axis).
template <typename T>
__global__ void quantize_kernel(const T* __restrict__ coords, int M, int D, int stride, double v0, double v1, double v2,
                                int div_f32, int* __restrict__ voxel, int* __restrict__ minmax) {
    int lo[3] = {0x7FFFFFFF, 0x7FFFFFFF, 0x7FFFFFFF}, hi[3] = {(int)0x80000000, (int)0x80000000, (int)0x80000000};
    const double vs[3] = {v0, v1, v2};
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
            if (d >= D) break;
            const T c = coords[(size_t)i * stride + d];
            double q;
            if (div_f32) q = (double)floorf(__fdiv_rn((float)c, (float)vs[d]));
            else q = floor((double)c / vs[d]);
            const int vi = (int)q;
            voxel[(size_t)i * D + d] = vi;
            lo[d] = min(lo[d], vi);
            hi[d] = max(hi[d], vi);
        }
    }
    __shared__ int s_lo[3], s_hi[3];
    if (threadIdx.x < 3) { s_lo[threadIdx.x] = 0x7FFFFFFF; s_hi[threadIdx.x] = (int)0x80000000; }
    __syncthreads();
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (d >= D) break;
        int a = lo[d], b = hi[d];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            a = min(a, __shfl_xor_sync(0xffffffffu, a, o));
            b = max(b, __shfl_xor_sync(0xffffffffu, b, o));
        }
        if ((threadIdx.x & 31) == 0) { atomicMin(&s_lo[d], a); atomicMax(&s_hi[d], b); }
    }
    __syncthreads();
    if (threadIdx.x < D) {
        atomicMin(&minmax[threadIdx.x], s_lo[threadIdx.x]);
        atomicMax(&minmax[3 + threadIdx.x], s_hi[threadIdx.x]);
    }
}

struct HashGeom {
    int D;
    int lo[3];
    unsigned long long ext[3];   // per-axis extent (max - min + 1)
};

__device__ __forceinline__ unsigned long long ravel_key(const int* v, const HashGeom& g) {
    // metric_utils.py:28-41: h = ((x0) * ext1 + x1) * ext2 + x2 on the min-shifted coordinates (uint64)
    unsigned long long h = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (d >= g.D) break;
        if (d > 0) h *= g.ext[d];
        h += (unsigned long long)((long long)v[d] - (long long)g.lo[d]);
    }
    return h;
}

__global__ void ravel_hash_kernel(const int* __restrict__ voxel, int M, HashGeom g, unsigned long long* __restrict__ out) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x)
        out[i] = ravel_key(voxel + (size_t)i * g.D, g);
}

__global__ void sq_set_bits_kernel(const int* __restrict__ voxel, int M, HashGeom g, unsigned* __restrict__ bitmap) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        const unsigned long long k = ravel_key(voxel + (size_t)i * g.D, g);
        atomicOr(bitmap + (k >> 5), 1u << (unsigned)(k & 31));
    }
}

constexpr int SQ_THREADS = 256;
constexpr int SQ_WPT = 8;                       // words per thread
constexpr int SQ_CHUNK = SQ_THREADS * SQ_WPT;   // bitmap words per block

// pass 1: set bits per chunk of 2048 words
__global__ void __launch_bounds__(SQ_THREADS) sq_chunk_popc_kernel(const unsigned* __restrict__ bitmap, size_t nwords,
                                                                   unsigned* __restrict__ chunk_sums) {
    const size_t base = (size_t)blockIdx.x * SQ_CHUNK;
    unsigned s = 0;
#pragma unroll
    for (int j = 0; j < SQ_WPT; ++j) {
        const size_t w = base + (size_t)j * SQ_THREADS + threadIdx.x;
        if (w < nwords) s += __popc(bitmap[w]);
    }
    __shared__ unsigned ws[SQ_THREADS / 32];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((threadIdx.x & 31) == 0) ws[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned t = 0;
        for (int w = 0; w < SQ_THREADS / 32; ++w) t += ws[w];
        chunk_sums[blockIdx.x] = t;
    }
}

// pass 2: exclusive scan of the chunk sums in place (one block), total -> n_unique
__global__ void __launch_bounds__(1024) sq_scan_chunks_kernel(unsigned* __restrict__ chunk_sums, int nchunks,
                                                              int* __restrict__ n_unique) {
    __shared__ unsigned part[1024];
    const int per = (nchunks + 1023) / 1024;
    const int lo = threadIdx.x * per, hi = min(lo + per, nchunks);
    unsigned s = 0;
    for (int i = lo; i < hi; ++i) s += chunk_sums[i];
    part[threadIdx.x] = s;
    __syncthreads();
    // Hillis-Steele inclusive scan over the 1024 partials
    for (int o = 1; o < 1024; o <<= 1) {
        const unsigned v = threadIdx.x >= o ? part[threadIdx.x - o] : 0u;
        __syncthreads();
        part[threadIdx.x] += v;
        __syncthreads();
    }
    unsigned run = threadIdx.x ? part[threadIdx.x - 1] : 0u;
    for (int i = lo; i < hi; ++i) {
        const unsigned c = chunk_sums[i];
        chunk_sums[i] = run;
        run += c;
    }
    if (threadIdx.x == 1023) *n_unique = (int)part[1023];
}

// pass 3: word_rank[w] = number of set bits in words < w
__global__ void __launch_bounds__(SQ_THREADS) sq_word_rank_kernel(const unsigned* __restrict__ bitmap, size_t nwords,
                                                                  const unsigned* __restrict__ chunk_sums,
                                                                  unsigned* __restrict__ word_rank) {
    // thread t owns the SQ_WPT CONSECUTIVE words base + t*SQ_WPT .. (an exclusive scan needs a linear order)
    const size_t base = (size_t)blockIdx.x * SQ_CHUNK + (size_t)threadIdx.x * SQ_WPT;
    unsigned c[SQ_WPT];
    unsigned s = 0;
#pragma unroll
    for (int j = 0; j < SQ_WPT; ++j) {
        c[j] = (base + j < nwords) ? __popc(bitmap[base + j]) : 0u;
        s += c[j];
    }
    // block-exclusive scan of s
    __shared__ unsigned wsum[SQ_THREADS / 32];
    unsigned incl = s;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned v = __shfl_up_sync(0xffffffffu, incl, o);
        if ((threadIdx.x & 31) >= o) incl += v;
    }
    if ((threadIdx.x & 31) == 31) wsum[threadIdx.x >> 5] = incl;
    __syncthreads();
    unsigned woff = 0;
    for (int w = 0; w < (int)(threadIdx.x >> 5); ++w) woff += wsum[w];
    unsigned run = chunk_sums[blockIdx.x] + woff + incl - s;
#pragma unroll
    for (int j = 0; j < SQ_WPT; ++j) {
        if (base + j < nwords) word_rank[base + j] = run;
        run += c[j];
    }
}

// pass 4: rank of every point = inverse index; first occurrence per rank
__global__ void sq_rank_points_kernel(const int* __restrict__ voxel, int M, HashGeom g, const unsigned* __restrict__ bitmap,
                                      const unsigned* __restrict__ word_rank, long long* __restrict__ inverse,
                                      int* __restrict__ first_idx) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < M; i += gridDim.x * blockDim.x) {
        const unsigned long long k = ravel_key(voxel + (size_t)i * g.D, g);
        const size_t w = (size_t)(k >> 5);
        const unsigned below = bitmap[w] & ((1u << (unsigned)(k & 31)) - 1u);
        const unsigned r = word_rank[w] + __popc(below);
        if (inverse) inverse[i] = (long long)r;
        atomicMin(first_idx + r, i);
    }
}

// pass 5: unique coordinates / first indices in hash order
__global__ void sq_emit_kernel(const int* __restrict__ voxel, int D, const int* __restrict__ first_idx,
                               const int* __restrict__ n_unique, int* __restrict__ uniq, long long* __restrict__ indices,
                               int M) {
    const int n = *n_unique;
    for (int r = blockIdx.x * blockDim.x + threadIdx.x; r < n && r < M; r += gridDim.x * blockDim.x) {
        const int idx = first_idx[r];
        if (indices) indices[r] = idx;
        if (uniq)
            for (int d = 0; d < D; ++d) uniq[(size_t)r * D + d] = voxel[(size_t)idx * D + d];
    }
}

// ---------------------------------------------------------------------------------------------------------
// occupancy grids with known bounds
// ---------------------------------------------------------------------------------------------------------
struct OccParams {
    const float* pcd;       // [total, stride]
    const int* offsets;     // [n_clouds + 1] (null: one cloud of M points)
    int M, stride, D;       // D = 2 (bev) or 3 (volume)
    float lo[3], hi[3];     // strict range test per axis
    float voxel;            // fp32(voxel_size)
    int minb[3], dims[3];
};

__device__ __forceinline__ bool occ_cell(const OccParams& p, const float* pt, size_t& cell) {
    size_t c = 0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
        if (d >= p.D) break;
        const float v = pt[d];
        if (!(v > p.lo[d] && v < p.hi[d])) return false;
        const int q = (int)floorf(__fdiv_rn(v, p.voxel)) - p.minb[d];
        if (q < 0 || q >= p.dims[d]) return false;   // fp32 rounding at the upper face (the reference would raise IndexError)
        c = c * (size_t)p.dims[d] + (size_t)q;
    }
    cell = c;
    return true;
}

// pcd2bev_sum: volume_sum[cell] += 1 once per (cloud, cell): per-cloud bitmap, the thread that sets a bit first counts
__global__ void bev_sum_kernel(const OccParams p, unsigned* __restrict__ bitmaps, size_t words_per_cloud,
                               float* __restrict__ volume_sum) {
    const int cl = blockIdx.y;
    const int lo = p.offsets ? p.offsets[cl] : 0, hi = p.offsets ? p.offsets[cl + 1] : p.M;
    unsigned* bm = bitmaps + (size_t)cl * words_per_cloud;
    for (int i = lo + blockIdx.x * blockDim.x + threadIdx.x; i < hi; i += gridDim.x * blockDim.x) {
        size_t cell;
        if (!occ_cell(p, p.pcd + (size_t)i * p.stride, cell)) continue;
        const unsigned bit = 1u << (unsigned)(cell & 31);
        const unsigned old = atomicOr(bm + (cell >> 5), bit);
        if (!(old & bit)) atomicAdd(volume_sum + cell, 1.0f);      // integer-valued fp32 adds: exact, order independent
    }
}

// pcd2voxel_full: vol[cell] = 1
__global__ void voxel_occ_kernel(const OccParams p, float* __restrict__ vol) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < p.M; i += gridDim.x * blockDim.x) {
        size_t cell;
        if (occ_cell(p, p.pcd + (size_t)i * p.stride, cell)) vol[cell] = 1.0f;
    }
}

static int grid_for(long long n, int threads, int cap = 148 * 16) {
    long long g = (n + threads - 1) / threads;
    if (g < 1) g = 1;
    return (int)(g > cap ? cap : g);
}

}  // namespace b200

using namespace b200;

extern "C" int b200_pcd2range(const float* pcd, const int* npts, const float* feature, float* proj_range,
                              float* proj_feature, void* zbuf, int F, int M, int H, int W, float fov_up_deg,
                              float fov_down_deg, float depth_min, float depth_max, float feature_fill, void* stream) {
    B200_CHECK_ARG(pcd && proj_range && zbuf && F > 0 && M >= 0 && H > 0 && W > 0 && F <= 65535);
    B200_CHECK_ARG((feature == nullptr) == (proj_feature == nullptr));
    cudaStream_t st = (cudaStream_t)stream;
    // fov_up = fov[0] / 180.0 * np.pi etc. are python floats (fp64); they meet fp32 arrays as fp32 scalars (NumPy >= 2)
    const double up = (double)fov_up_deg / 180.0 * 3.14159265358979323846;
    const double down = (double)fov_down_deg / 180.0 * 3.14159265358979323846;
    P2RParams p{pcd, npts, feature, proj_range, proj_feature, (unsigned long long*)zbuf, F, M, H, W,
                (float)fabs(down), (float)(fabs(down) + fabs(up)), depth_min, depth_max, feature_fill};
    if (cudaMemsetAsync(zbuf, 0xFF, (size_t)F * H * W * 8, st) != cudaSuccess) {
        set_error("pcd2range: memset failed");
        return B200_E_CUDA;
    }
    if (M > 0) {
        p2r_scatter_kernel<<<dim3(grid_for(M, 256), F), 256, 0, st>>>(p);
        B200_CHECK_LAUNCH();
    }
    p2r_gather_kernel<<<dim3(grid_for((long long)H * W, 256), F), 256, 0, st>>>(p);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

extern "C" int b200_range2xyz(const float* range_img, double* xyz, int F, int H, int W, float fov_up_deg,
                              float fov_down_deg, float depth_min, float depth_max, float depth_scale, int log_scale,
                              void* stream) {
    B200_CHECK_ARG(range_img && xyz && F > 0 && H > 0 && W > 0 && F <= 65535);
    const double up = (double)fov_up_deg / 180.0 * 3.14159265358979323846;
    const double down = (double)fov_down_deg / 180.0 * 3.14159265358979323846;
    range2xyz_kernel<<<dim3(grid_for((long long)H * W, 256), F), 256, 0, (cudaStream_t)stream>>>(
        range_img, xyz, H, W, fabs(down), fabs(down) + fabs(up), depth_min, depth_max, depth_scale, log_scale);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

extern "C" int b200_quantize_coords(const void* coords, int coords_f64, int M, int D, int stride, double v0, double v1,
                                    double v2, int div_f32, int32_t* voxel, int32_t* minmax, void* stream) {
    B200_CHECK_ARG(coords && voxel && minmax && M > 0 && (D == 2 || D == 3) && stride >= D);
    B200_CHECK_ARG(v0 > 0 && v1 > 0 && (D == 2 || v2 > 0));
    cudaStream_t st = (cudaStream_t)stream;
    minmax_init_kernel<<<1, 32, 0, st>>>(minmax);
    B200_CHECK_LAUNCH();
    const int g = grid_for(M, 256, 148 * 8);
    if (coords_f64)
        quantize_kernel<double><<<g, 256, 0, st>>>((const double*)coords, M, D, stride, v0, v1, v2, div_f32, voxel, minmax);
    else
        quantize_kernel<float><<<g, 256, 0, st>>>((const float*)coords, M, D, stride, v0, v1, v2, div_f32, voxel, minmax);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

static int make_geom(const int32_t* minmax_host, int D, HashGeom& g, unsigned long long& nkeys) {
    g.D = D;
    nkeys = 1;
    for (int d = 0; d < 3; ++d) {
        g.lo[d] = d < D ? minmax_host[d] : 0;
        const long long e = d < D ? (long long)minmax_host[3 + d] - (long long)minmax_host[d] + 1 : 1;
        if (e <= 0) return -1;
        g.ext[d] = (unsigned long long)e;
        if (nkeys > (1ull << 40) / g.ext[d]) return -1;   // absurd extents (outlier points): refuse instead of overflowing
        nkeys *= g.ext[d];
    }
    return 0;
}

static const unsigned long long SQ_MAX_KEYS = 1ull << 35;   // 4 GiB bitmap + 4 GiB word ranks at most

extern "C" size_t b200_sparse_quantize_workspace(const int32_t* minmax_host, int D, int M) {
    HashGeom g;
    unsigned long long nkeys;
    if (!minmax_host || M <= 0 || (D != 2 && D != 3) || make_geom(minmax_host, D, g, nkeys) || nkeys > SQ_MAX_KEYS) return 0;
    const size_t nwords = (size_t)((nkeys + 31) / 32);
    const size_t nchunks = (nwords + SQ_CHUNK - 1) / SQ_CHUNK;
    // bitmap | word_rank | chunk_sums | first_idx   (each rounded up to 256 bytes)
    auto r256 = [](size_t b) { return (b + 255) / 256 * 256; };
    return r256(nwords * 4) + r256(nwords * 4) + r256(nchunks * 4) + r256((size_t)M * 4);
}

extern "C" int b200_ravel_hash(const int32_t* voxel, int M, int D, const int32_t* minmax_host, uint64_t* out, void* stream) {
    B200_CHECK_ARG(voxel && out && minmax_host && M > 0 && (D == 2 || D == 3));
    HashGeom g;
    unsigned long long nkeys;
    B200_CHECK_ARG(make_geom(minmax_host, D, g, nkeys) == 0);
    ravel_hash_kernel<<<grid_for(M, 256), 256, 0, (cudaStream_t)stream>>>(voxel, M, g, (unsigned long long*)out);
    B200_CHECK_LAUNCH();
    return B200_OK;
}

extern "C" int b200_sparse_quantize(const int32_t* voxel, int M, int D, const int32_t* minmax_host, void* workspace,
                                    size_t workspace_bytes, int32_t* uniq_coords, int64_t* indices, int64_t* inverse,
                                    int32_t* n_unique, void* stream) {
    B200_CHECK_ARG(voxel && minmax_host && workspace && n_unique && M > 0 && (D == 2 || D == 3));
    HashGeom g;
    unsigned long long nkeys;
    if (make_geom(minmax_host, D, g, nkeys) || nkeys > SQ_MAX_KEYS) {
        set_error("sparse_quantize: bounding grid of %llu cells exceeds the %llu-cell bitmap limit", nkeys, SQ_MAX_KEYS);
        return B200_E_ARG;
    }
    const size_t need = b200_sparse_quantize_workspace(minmax_host, D, M);
    B200_CHECK_ARG(need > 0 && workspace_bytes >= need);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t nwords = (size_t)((nkeys + 31) / 32);
    const size_t nchunks = (nwords + SQ_CHUNK - 1) / SQ_CHUNK;
    B200_CHECK_ARG(nchunks <= 0x7FFFFFFFull);
    auto r256 = [](size_t b) { return (b + 255) / 256 * 256; };
    uint8_t* ws = (uint8_t*)workspace;
    unsigned* bitmap = (unsigned*)ws;
    unsigned* word_rank = (unsigned*)(ws + r256(nwords * 4));
    unsigned* chunk_sums = (unsigned*)(ws + 2 * r256(nwords * 4));
    int* first_idx = (int*)(ws + 2 * r256(nwords * 4) + r256(nchunks * 4));
    if (cudaMemsetAsync(bitmap, 0, nwords * 4, st) != cudaSuccess ||
        cudaMemsetAsync(first_idx, 0x7F, (size_t)M * 4, st) != cudaSuccess) {
        set_error("sparse_quantize: memset failed");
        return B200_E_CUDA;
    }
    const int gp = grid_for(M, 256);
    sq_set_bits_kernel<<<gp, 256, 0, st>>>(voxel, M, g, bitmap);
    B200_CHECK_LAUNCH();
    sq_chunk_popc_kernel<<<(unsigned)nchunks, SQ_THREADS, 0, st>>>(bitmap, nwords, chunk_sums);
    B200_CHECK_LAUNCH();
    sq_scan_chunks_kernel<<<1, 1024, 0, st>>>(chunk_sums, (int)nchunks, n_unique);
    B200_CHECK_LAUNCH();
    sq_word_rank_kernel<<<(unsigned)nchunks, SQ_THREADS, 0, st>>>(bitmap, nwords, chunk_sums, word_rank);
    B200_CHECK_LAUNCH();
    sq_rank_points_kernel<<<gp, 256, 0, st>>>(voxel, M, g, bitmap, word_rank, (long long*)inverse, first_idx);
    B200_CHECK_LAUNCH();
    if (uniq_coords || indices) {
        sq_emit_kernel<<<gp, 256, 0, st>>>(voxel, D, first_idx, n_unique, uniq_coords, (long long*)indices, M);
        B200_CHECK_LAUNCH();
    }
    return B200_OK;
}

static int fill_occ(OccParams& p, const float* pcd, const int* offsets, int M, int stride, int D, const float* lo,
                    const float* hi, float voxel, const int* minb, const int* dims) {
    p.pcd = pcd; p.offsets = offsets; p.M = M; p.stride = stride; p.D = D; p.voxel = voxel;
    for (int d = 0; d < 3; ++d) {
        p.lo[d] = d < D ? lo[d] : 0.f; p.hi[d] = d < D ? hi[d] : 0.f;
        p.minb[d] = d < D ? minb[d] : 0; p.dims[d] = d < D ? dims[d] : 1;
        if (d < D && dims[d] <= 0) return -1;
    }
    return 0;
}

extern "C" int b200_bev_occupancy_sum(const float* pcd, const int* offsets, int n_clouds, int max_cloud_pts, int stride,
                                      float x_lo, float x_hi, float y_lo, float y_hi, float voxel, int min_bx, int min_by,
                                      int X, int Y, void* bitmap_ws, float* volume_sum, void* stream) {
    B200_CHECK_ARG(pcd && offsets && bitmap_ws && volume_sum && n_clouds > 0 && n_clouds <= 65535 && stride >= 2);
    B200_CHECK_ARG(voxel > 0.f && max_cloud_pts >= 0);
    const float lo[2] = {x_lo, y_lo}, hi[2] = {x_hi, y_hi};
    const int minb[2] = {min_bx, min_by}, dims[2] = {X, Y};
    OccParams p;
    B200_CHECK_ARG(fill_occ(p, pcd, offsets, 0, stride, 2, lo, hi, voxel, minb, dims) == 0);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t words = ((size_t)X * Y + 31) / 32;
    if (cudaMemsetAsync(bitmap_ws, 0, words * 4 * (size_t)n_clouds, st) != cudaSuccess) {
        set_error("bev_occupancy_sum: memset failed");
        return B200_E_CUDA;
    }
    if (max_cloud_pts > 0) {
        bev_sum_kernel<<<dim3(grid_for(max_cloud_pts, 256, 64), n_clouds), 256, 0, st>>>(p, (unsigned*)bitmap_ws, words,
                                                                                         volume_sum);
        B200_CHECK_LAUNCH();
    }
    return B200_OK;
}

extern "C" int b200_voxel_occupancy(const float* pcd, int M, int stride, const float* range_lo_hi_host, float voxel,
                                    const int32_t* min_bound_host, const int32_t* dims_host, float* vol, void* stream) {
    B200_CHECK_ARG(pcd && vol && range_lo_hi_host && min_bound_host && dims_host && M >= 0 && stride >= 3 && voxel > 0.f);
    const float lo[3] = {range_lo_hi_host[0], range_lo_hi_host[2], range_lo_hi_host[4]};
    const float hi[3] = {range_lo_hi_host[1], range_lo_hi_host[3], range_lo_hi_host[5]};
    OccParams p;
    B200_CHECK_ARG(fill_occ(p, pcd, nullptr, M, stride, 3, lo, hi, voxel, min_bound_host, dims_host) == 0);
    cudaStream_t st = (cudaStream_t)stream;
    const size_t cells = (size_t)dims_host[0] * dims_host[1] * dims_host[2];
    if (cudaMemsetAsync(vol, 0, cells * 4, st) != cudaSuccess) {
        set_error("voxel_occupancy: memset failed");
        return B200_E_CUDA;
    }
    if (M > 0) {
        voxel_occ_kernel<<<grid_for(M, 256), 256, 0, st>>>(p, vol);
        B200_CHECK_LAUNCH();
    }
    return B200_OK;
}
