/*
 * oracle/lidar_ops.c -- plain-C CPU restatement of the point-cloud side of the hot path.  TEST ORACLE ONLY.
 *
 *   oracle_range_project        <- lidargen/dataset/transforms_3d/common.py:26-91 (load_points_as_images,
 *                                   scan_unfolding=False), NumPy >= 2 promotion rules (grid_h path in fp64)
 *   oracle_points_in_boxes      <- lidargen/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:121-168 (MARGIN 1e-2)
 *   oracle_points_in_boxes_first<- lidargen/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu:23-36,313-336
 *   oracle_voxel_index          <- lidargen/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu:39-75
 *
 * Definitions that make the integer outputs platform independent (the reference itself is not: it depends on
 * the libm / SVML flavour NumPy was built with and on nvcc's FMA contraction):
 *   - fp32 asin/atan2/cos/sin := (float) f((double) x)   (correctly rounded up to double rounding)
 *   - no FMA contraction (compile with -ffp-contract=off); every fp32 op rounds to nearest
 *   - z-buffer ties (identical depth in one pixel): the point with the highest index wins
 * Pinning: tests/golden/projection.npz holds the outputs of the reference's own NumPy function on a seeded
 * synthetic sweep (tests/golden/make_golden_lidar.py); points_in_boxes / voxel_index have no runnable
 * reference in this container (torch extension, un-buildable) -> "parity unpinned" for those two.
 */
#include <math.h>
#include <stdint.h>
#include <string.h>

static float asin_f(float x) { return (float)asin((double)x); }
static float atan2_f(float y, float x) { return (float)atan2((double)y, (double)x); }
static float cos_f(float x) { return (float)cos((double)x); }
static float sin_f(float x) { return (float)sin((double)x); }

void oracle_project_point(float x, float y, float z, int H, int W, double h_up, double h_down, float* depth_out,
                          int* gh_out, int* gw_out) {
    volatile float xx = x * x, yy = y * y, zz = z * z;
    volatile float s = xx + yy;
    s = s + zz;
    float depth = sqrtf(s);
    volatile float den = depth + 1e-6f;
    float ratio = z / den;
    double elev = (double)asin_f(ratio) + fabs(h_down);
    double g = 1.0 - elev / (h_up - h_down);
    g = floor(g * (double)H);
    if (g < 0.0) g = 0.0;
    if (g > (double)(H - 1)) g = (double)(H - 1);
    float az = -atan2_f(y, x);
    volatile float t = az / 3.14159274101257324f;
    t = t + 1.0f;
    t = t * 0.5f;
    float tm = fmodf(t, 1.0f);
    if (tm < 0.f) tm += 1.0f;
    volatile float gwv = tm * (float)W;
    float gwf = floorf(gwv);
    if (gwf < 0.f) gwf = 0.f;
    if (gwf > (float)(W - 1)) gwf = (float)(W - 1);
    *depth_out = depth;
    *gh_out = (int)g;
    *gw_out = (int)gwf;
}

/* points [M,4]; out [H,W,6] (x,y,z,i,depth,mask); grid [M,2]; winner [H,W] (index of the surviving point, -1) */
void oracle_range_project(const float* pts, int M, int H, int W, float min_d, float max_d, float fov_up_deg,
                          float fov_down_deg, float* out, int* grid, int* winner) {
    const double d2r = 3.14159265358979323846 / 180.0;
    const double h_up = (double)fov_up_deg * d2r, h_down = (double)fov_down_deg * d2r;
    memset(out, 0, sizeof(float) * (size_t)H * W * 6);
    for (int i = 0; i < H * W; ++i) winner[i] = -1;
    for (int i = 0; i < M; ++i) {
        float depth;
        int gh, gw;
        oracle_project_point(pts[i * 4], pts[i * 4 + 1], pts[i * 4 + 2], H, W, h_up, h_down, &depth, &gh, &gw);
        if (grid) { grid[i * 2] = gh; grid[i * 2 + 1] = gw; }
        const int p = gh * W + gw;
        const int cur = winner[p];
        if (cur < 0 || depth <= out[p * 6 + 4]) {  /* nearer wins; equal depth: later (higher) index wins */
            winner[p] = i;
            out[p * 6 + 0] = pts[i * 4]; out[p * 6 + 1] = pts[i * 4 + 1]; out[p * 6 + 2] = pts[i * 4 + 2];
            out[p * 6 + 3] = pts[i * 4 + 3];
            out[p * 6 + 4] = depth;
            out[p * 6 + 5] = (depth >= min_d && depth <= max_d) ? 1.f : 0.f;
        }
    }
}

static int pt_in_box(const float* pt, const float* bx, float margin, float* lx, float* ly) {
    const float x = pt[0], y = pt[1], z = pt[2];
    const float cx = bx[0], cy = bx[1], cz = bx[2], dx = bx[3], dy = bx[4], dz = bx[5], rz = bx[6];
    volatile float zc = z - cz;
    if ((double)fabsf(zc) > (double)dz / 2.0) return 0;
    volatile float sx = x - cx, sy = y - cy;
    const float cosa = cos_f(-rz), sina = sin_f(-rz);
    volatile float a = sx * cosa, b = sy * (-sina), c = sx * sina, d = sy * cosa;
    volatile float llx = a + b, lly = c + d;
    *lx = llx; *ly = lly;
    return ((double)fabsf(llx) < (double)dx / 2.0 + (double)margin) &&
           ((double)fabsf(lly) < (double)dy / 2.0 + (double)margin);
}

void oracle_points_in_boxes(const float* pts, const float* boxes, int N, int M, int* out) {
    float lx, ly;
    for (int i = 0; i < N; ++i)
        for (int j = 0; j < M; ++j) out[(size_t)i * M + j] = pt_in_box(pts + j * 3, boxes + i * 7, 1e-2f, &lx, &ly);
}

void oracle_points_in_boxes_first(const float* pts, const float* boxes, int B, int N, int M, int* out) {
    float lx, ly;
    for (int b = 0; b < B; ++b)
        for (int j = 0; j < M; ++j) {
            int r = -1;
            for (int k = 0; k < N; ++k)
                if (pt_in_box(pts + ((size_t)b * M + j) * 3, boxes + ((size_t)b * N + k) * 7, 1e-5f, &lx, &ly)) { r = k; break; }
            out[(size_t)b * M + j] = r;
        }
}

static unsigned clampu(unsigned v, unsigned hi) { return v > hi ? hi : v; }

void oracle_voxel_index(const float* pts, const float* rois, int N, int M, int ox, int oy, int oz, int* out) {
    for (int i = 0; i < N; ++i) {
        const float* bx = rois + i * 7;
        for (int j = 0; j < M; ++j) {
            float lx = 0.f, ly = 0.f;
            int code = -1;
            if (pt_in_box(pts + j * 3, bx, 1e-5f, &lx, &ly)) {
                volatile float lz = pts[j * 3 + 2] - bx[2];
                const float dx = bx[3], dy = bx[4], dz = bx[5];
                volatile float xr = dx / (float)ox, yr = dy / (float)oy, zr = dz / (float)oz;
                volatile float hx = dx / 2.f, hy = dy / 2.f, hz = dz / 2.f;
                volatile float ax = lx + hx, ay = ly + hy, az = lz + hz;
                volatile float qx = ax / xr, qy = ay / yr, qz = az / zr;
                /* unsigned idx = int(f); min(max(idx, 0), n-1) with the unsigned overloads of the reference */
                unsigned xi = clampu((unsigned)(int)qx, (unsigned)(ox - 1));
                unsigned yi = clampu((unsigned)(int)qy, (unsigned)(oy - 1));
                unsigned zi = clampu((unsigned)(int)qz, (unsigned)(oz - 1));
                code = (int)((xi << 16) + (yi << 8) + zi);
            }
            out[(size_t)i * M + j] = code;
        }
    }
}
