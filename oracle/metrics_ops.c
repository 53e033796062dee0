/*
 * oracle/metrics_ops.c -- plain-C CPU restatement of the evaluation-side point-cloud ops.  TEST ORACLE ONLY.
 *
 *   oracle_pcd2range        <- lidargen/metrics/metric_utils.py:65-121 (float32 input, NumPy >= 2 promotion: python
 *                              floats are weak, so every operation on the points stays fp32)
 *   oracle_range2xyz        <- metric_utils.py:124-154 (fp64 ray directions, fp32 depth)
 *   oracle_quantize         <- np.floor(coords / voxel_size).astype(np.int32)  (:51 fp64 division; :189,249,277,301 fp32)
 *   oracle_sparse_quantize  <- ravel_hash + np.unique(return_index, return_inverse) (:28-41,53-62), by SORTING
 *                              (key, index) pairs -- deliberately not the bitmap/popcount scheme of the CUDA path
 *   oracle_bev_sum          <- pcd2bev_sum (:231-256);  oracle_voxel_full <- pcd2voxel_full (:170-199)
 *
 * Same platform-independence definitions as oracle/lidar_ops.c: fp32 asin/atan2 := (float) f((double) x), no FMA
 * contraction (-ffp-contract=off), depth ties in a pixel -> lowest index wins (stable descending-depth order).
 * Pinning: tests/golden/metrics.npz holds the outputs of the reference's own functions on seeded sweeps
 * (tests/golden/make_golden_metrics.py).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static float asin_f(float x) { return (float)asin((double)x); }
static float atan2_f(float y, float x) { return (float)atan2((double)y, (double)x); }
static const double PI = 3.14159265358979323846;

/* pts [M,3]; feature [M] or NULL; proj_range [H,W]; proj_feature [H,W] or NULL; winner [H,W] (index or -1) */
void oracle_pcd2range(const float* pts, const float* feature, int M, int H, int W, float fov_up_deg, float fov_down_deg,
                      float dmin, float dmax, float feature_fill, float* proj_range, float* proj_feature, int* winner) {
    const double up = (double)fov_up_deg / 180.0 * PI, down = (double)fov_down_deg / 180.0 * PI;
    const float fov_down_abs = (float)fabs(down), fov_range = (float)(fabs(down) + fabs(up)), pi_f = (float)PI;
    for (int i = 0; i < H * W; ++i) {
        proj_range[i] = -1.f;
        winner[i] = -1;
        if (proj_feature) proj_feature[i] = feature_fill;
    }
    for (int i = 0; i < M; ++i) {
        const float x = pts[i * 3], y = pts[i * 3 + 1], z = pts[i * 3 + 2];
        volatile float xx = x * x, yy = y * y, zz = z * z;
        volatile float s = xx + yy;
        s = s + zz;
        const float depth = sqrtf(s);
        if (!(depth > dmin && depth < dmax)) continue;
        const float yaw = -atan2_f(y, x);
        volatile float ratio = z / depth;
        const float pitch = asin_f(ratio);
        volatile float px = yaw / pi_f;
        px = px + 1.0f;
        px = 0.5f * px;
        volatile float py = pitch + fov_down_abs;
        py = py / fov_range;
        py = 1.0f - py;
        px = px * (float)W;
        py = py * (float)H;
        float fx = floorf(px), fy = floorf(py);
        if (fx > (float)(W - 1)) fx = (float)(W - 1);
        if (fx < 0.f) fx = 0.f;
        if (fy > (float)(H - 1)) fy = (float)(H - 1);
        if (fy < 0.f) fy = 0.f;
        const int p = (int)fy * W + (int)fx;
        if (winner[p] < 0 || depth < proj_range[p]) {   /* strictly nearer replaces; ties keep the lower index */
            winner[p] = i;
            proj_range[p] = depth;
            if (proj_feature) proj_feature[p] = feature[i];
        }
    }
}

void oracle_range2xyz(const float* img, int H, int W, float fov_up_deg, float fov_down_deg, float dmin, float dmax,
                      float depth_scale, int log_scale, double* xyz) {
    const double up = (double)fov_up_deg / 180.0 * PI, down = (double)fov_down_deg / 180.0 * PI;
    const double fov_range = fabs(down) + fabs(up);
    const int hw = H * W;
    for (int r = 0; r < H; ++r)
        for (int c = 0; c < W; ++c) {
            const int i = r * W + c;
            float depth = img[i];
            if (log_scale) {
                volatile float e = img[i] * depth_scale;
                depth = exp2f(e) - 1.f;
            }
            const double yaw = PI * (((double)c / (double)W) * 2.0 - 1.0);
            const double pitch = (1.0 - (double)r / (double)H) * fov_range - fabs(down);
            const int ok = depth > dmin && depth < dmax;
            xyz[i] = ok ? cos(yaw) * cos(pitch) * (double)depth : -1.0;
            xyz[hw + i] = ok ? -sin(yaw) * cos(pitch) * (double)depth : -1.0;
            xyz[2 * hw + i] = ok ? sin(pitch) * (double)depth : -1.0;
        }
}

/* coords [M,stride] (fp32 or fp64), first D columns -> voxel int32 [M,D] */
void oracle_quantize(const void* coords, int is_f64, int M, int D, int stride, const double* vs, int div_f32, int* voxel) {
    for (int i = 0; i < M; ++i)
        for (int d = 0; d < D; ++d) {
            const double c = is_f64 ? ((const double*)coords)[(size_t)i * stride + d]
                                    : (double)((const float*)coords)[(size_t)i * stride + d];
            double q;
            if (div_f32) {
                volatile float t = (float)c / (float)vs[d];
                q = (double)floorf(t);
            } else {
                q = floor(c / vs[d]);
            }
            voxel[(size_t)i * D + d] = (int)q;
        }
}

typedef struct { uint64_t key; int idx; } KeyIdx;
static int cmp_keyidx(const void* a, const void* b) {
    const KeyIdx* x = (const KeyIdx*)a; const KeyIdx* y = (const KeyIdx*)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->idx < y->idx ? -1 : (x->idx > y->idx);
}

/* keys [M] out (ravel hash); uniq [M,D], indices [M], inverse [M]; returns the number of unique voxels */
int oracle_sparse_quantize(const int* voxel, int M, int D, uint64_t* keys, int* uniq, int64_t* indices, int64_t* inverse) {
    int lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
    for (int d = 0; d < D; ++d) {
        lo[d] = hi[d] = voxel[d];
        for (int i = 1; i < M; ++i) {
            const int v = voxel[(size_t)i * D + d];
            if (v < lo[d]) lo[d] = v;
            if (v > hi[d]) hi[d] = v;
        }
    }
    KeyIdx* ki = (KeyIdx*)malloc(sizeof(KeyIdx) * (size_t)M);
    for (int i = 0; i < M; ++i) {
        uint64_t h = 0;
        for (int k = 0; k < D - 1; ++k) {       /* metric_utils.py:36-39 */
            h += (uint64_t)((int64_t)voxel[(size_t)i * D + k] - lo[k]);
            h *= (uint64_t)((int64_t)hi[k + 1] - lo[k + 1] + 1);
        }
        h += (uint64_t)((int64_t)voxel[(size_t)i * D + D - 1] - lo[D - 1]);
        keys[i] = h;
        ki[i].key = h;
        ki[i].idx = i;
    }
    qsort(ki, (size_t)M, sizeof(KeyIdx), cmp_keyidx);
    int n = 0;
    for (int j = 0; j < M; ++j) {
        if (j == 0 || ki[j].key != ki[j - 1].key) {
            indices[n] = ki[j].idx;              /* smallest index of the run: first occurrence */
            for (int d = 0; d < D; ++d) uniq[(size_t)n * D + d] = voxel[(size_t)ki[j].idx * D + d];
            ++n;
        }
        inverse[ki[j].idx] = n - 1;
    }
    free(ki);
    return n;
}

static int occ_cell(const float* pt, int D, const float* lo, const float* hi, float voxel, const int* minb, const int* dims,
                    size_t* cell) {
    size_t c = 0;
    for (int d = 0; d < D; ++d) {
        if (!(pt[d] > lo[d] && pt[d] < hi[d])) return 0;
        volatile float t = pt[d] / voxel;
        const int q = (int)floorf(t) - minb[d];
        if (q < 0 || q >= dims[d]) return 0;
        c = c * (size_t)dims[d] + (size_t)q;
    }
    *cell = c;
    return 1;
}

/* clouds concatenated [total,stride], offsets [n+1]; volume_sum [X,Y] += 1 per (cloud, occupied cell) */
void oracle_bev_sum(const float* pcd, const int* offsets, int n_clouds, int stride, const float* lo, const float* hi,
                    float voxel, const int* minb, const int* dims, float* volume_sum) {
    const size_t cells = (size_t)dims[0] * dims[1];
    int* stamp = (int*)calloc(cells, sizeof(int));
    for (int c = 0; c < n_clouds; ++c)
        for (int i = offsets[c]; i < offsets[c + 1]; ++i) {
            size_t cell;
            if (!occ_cell(pcd + (size_t)i * stride, 2, lo, hi, voxel, minb, dims, &cell)) continue;
            if (stamp[cell] != c + 1) { stamp[cell] = c + 1; volume_sum[cell] += 1.f; }
        }
    free(stamp);
}

void oracle_voxel_full(const float* pcd, int M, int stride, const float* lo, const float* hi, float voxel, const int* minb,
                       const int* dims, float* vol) {
    memset(vol, 0, sizeof(float) * (size_t)dims[0] * dims[1] * dims[2]);
    for (int i = 0; i < M; ++i) {
        size_t cell;
        if (occ_cell(pcd + (size_t)i * stride, 3, lo, hi, voxel, minb, dims, &cell)) vol[cell] = 1.f;
    }
}
