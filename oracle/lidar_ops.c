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
#include <stdlib.h>
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

/* ---------------------------------------------------------------------------------------------------------
 * oracle_range_project_f64 <- the same load_points_as_images on a float64 point array (what the temporal glue feeds it:
 * tools/vis_tools/utils/pipe_related.py:244-255 warps the background with a float64 4x4 and re-projects it; every
 * intermediate is float64 and only the final image is cast to float32, common.py:91).
 * Pinned by tests/golden/temporal.npz (next-frame point sets of the reference's own get_next_frame_points).
 * --------------------------------------------------------------------------------------------------------- */
void oracle_range_project_f64(const double* pts, int M, int H, int W, double min_d, double max_d, double fov_up_deg,
                              double fov_down_deg, float* out, int* winner) {
    const double d2r = 3.14159265358979323846 / 180.0;
    const double h_up = fov_up_deg * d2r, h_down = fov_down_deg * d2r;
    double* best = (double*)malloc(sizeof(double) * (size_t)H * W);
    memset(out, 0, sizeof(float) * (size_t)H * W * 6);
    for (int i = 0; i < H * W; ++i) winner[i] = -1;
    for (int i = 0; i < M; ++i) {
        const double x = pts[i * 4], y = pts[i * 4 + 1], z = pts[i * 4 + 2];
        volatile double xx = x * x, yy = y * y, zz = z * z;
        volatile double s = xx + yy;
        s = s + zz;
        const double depth = sqrt(s);
        volatile double den = depth + 1e-6;
        volatile double ratio = z / den;
        volatile double q = (asin(ratio) + fabs(h_down)) / (h_up - h_down);
        volatile double g1 = 1.0 - q;
        volatile double g2 = g1 * (double)H;
        double g = floor(g2);
        if (g < 0.0) g = 0.0;
        if (g > (double)(H - 1)) g = (double)(H - 1);
        volatile double t = -atan2(y, x) / 3.141592653589793;
        t = t + 1.0;
        t = t * 0.5;
        double tm = fmod(t, 1.0);
        if (tm < 0.0) tm += 1.0;
        volatile double gwv = tm * (double)W;
        double gw = floor(gwv);
        if (gw < 0.0) gw = 0.0;
        if (gw > (double)(W - 1)) gw = (double)(W - 1);
        const int p = (int)g * W + (int)gw;
        if (winner[p] < 0 || depth <= best[p]) {   /* nearer wins; equal depth: later (higher) index wins */
            winner[p] = i;
            best[p] = depth;
            out[p * 6 + 0] = (float)x; out[p * 6 + 1] = (float)y; out[p * 6 + 2] = (float)z;
            out[p * 6 + 3] = (float)pts[i * 4 + 3];
            out[p * 6 + 4] = (float)depth;
            out[p * 6 + 5] = (depth >= min_d && depth <= max_d) ? 1.f : 0.f;
        }
    }
    free(best);
}

/* ---------------------------------------------------------------------------------------------------------
 * oracle_boxes_to_mask <- lidargen/dataset/transforms_3d/common.py:99-216 (convert_boxes_to_2d + convert_points_to_2d).
 * boxes [N,8] (x, y, z, l, w, h, yaw, class) given as float32 (f64 = 0) or float64 (f64 = 1); the reference evaluates
 * the corner offsets, cos / sin of the yaw and the centre depth in the array's dtype and everything after in float64.
 * boxes_2d [N,4] float64 (x1, y1, x2, y2 normalised), mask [2,H,W] float32, weight [H,W] float32.
 * Pinned by tests/golden/boxes2d.npz and temporal.npz (goldens of the unmodified function, both dtypes).
 * --------------------------------------------------------------------------------------------------------- */
void oracle_boxes_to_mask(const void* boxes, int f64, int N, int H, int W, double fov_up_deg, double fov_down_deg,
                          double* boxes_2d, float* mask, float* weight) {
    const double d2r = 3.14159265358979323846 / 180.0;
    const double h_up = fov_up_deg * d2r, h_down = fov_down_deg * d2r;
    static const int SX[8] = {1, 1, -1, -1, 1, 1, -1, -1}, SY[8] = {1, -1, -1, 1, 1, -1, -1, 1}, SZ[8] = {1, 1, 1, 1, -1, -1, -1, -1};
    int* rect = (int*)malloc(sizeof(int) * 5 * (size_t)N);
    float* ow = (float*)malloc(sizeof(float) * (size_t)N);
    float* cdep = (float*)malloc(sizeof(float) * (size_t)N);
    float* cls = (float*)malloc(sizeof(float) * (size_t)N);
    int max_area = -2147483647;
    memset(mask, 0, sizeof(float) * 2 * (size_t)H * W);
    for (int i = 0; i < N; ++i) {
        double b[8], hx, hy, hz, c, s;
        if (f64) {
            const double* bd = (const double*)boxes + i * 8;
            for (int k = 0; k < 8; ++k) b[k] = bd[k];
            hx = b[3] / 2; hy = b[4] / 2; hz = b[5] / 2;
            c = cos(b[6]); s = sin(b[6]);
            volatile double xx = b[0] * b[0], yy = b[1] * b[1], zz = b[2] * b[2];
            volatile double sm = xx + yy;
            sm = sm + zz;
            volatile double cd = sqrt(sm) + 1e-6;
            cdep[i] = (float)cd;
        } else {
            const float* bf = (const float*)boxes + i * 8;
            for (int k = 0; k < 8; ++k) b[k] = (double)bf[k];
            volatile float fx = bf[3] / 2.f, fy = bf[4] / 2.f, fz = bf[5] / 2.f;
            hx = fx; hy = fy; hz = fz;
            c = (double)cos_f(bf[6]); s = (double)sin_f(bf[6]);
            volatile float xx = bf[0] * bf[0], yy = bf[1] * bf[1], zz = bf[2] * bf[2];
            volatile float sm = xx + yy;
            sm = sm + zz;
            volatile float cd = sqrtf(sm) + 1e-6f;
            cdep[i] = cd;
        }
        cls[i] = (float)b[7];
        double u0 = 0, u1 = 0, v0 = 0, v1 = 0;
        for (int k = 0; k < 8; ++k) {
            const double lx = SX[k] * hx, ly = SY[k] * hy, lz = SZ[k] * hz;
            volatile double a1 = c * lx, a2 = -s * ly, a3 = s * lx, a4 = c * ly;
            volatile double rx = a1 + a2, ry = a3 + a4;
            volatile double x = rx + b[0], y = ry + b[1], z = lz + b[2];
            volatile double xx = x * x, yy = y * y, zz = z * z;
            volatile double sm = xx + yy;
            sm = sm + zz;
            volatile double depth = sqrt(sm) + 1e-6;
            volatile double ratio = z / depth;
            volatile double q = (asin(ratio) + fabs(h_down)) / (h_up - h_down);
            volatile double g1 = 1.0 - q;
            volatile double g2 = g1 * (double)H;
            double g = floor(g2);
            if (g < 0.0) g = 0.0;
            if (g > (double)(H - 1)) g = (double)(H - 1);
            volatile double t = -atan2(y, x) / 3.141592653589793;
            t = t + 1.0;
            t = t * 0.5;
            double tm = fmod(t, 1.0);
            if (tm < 0.0) tm += 1.0;
            volatile double gwv = tm * (double)W;
            double gw = floor(gwv);
            if (gw < 0.0) gw = 0.0;
            if (gw > (double)(W - 1)) gw = (double)(W - 1);
            volatile double u = gw / (double)W, v = g / (double)H;
            if (k == 0) { u0 = u1 = u; v0 = v1 = v; }
            else {
                if (u < u0) u0 = u;
                if (u > u1) u1 = u;
                if (v < v0) v0 = v;
                if (v > v1) v1 = v;
            }
        }
        boxes_2d[i * 4] = u0; boxes_2d[i * 4 + 1] = v0; boxes_2d[i * 4 + 2] = u1; boxes_2d[i * 4 + 3] = v1;
        volatile double px1 = u0 * W, px2 = u1 * W, py1 = v0 * H, py2 = v1 * H;
        int* r = rect + i * 5;
        r[0] = (int)px1; r[1] = (int)py1; r[2] = (int)px2; r[3] = (int)py2;
        r[4] = ((double)(r[2] - r[0]) / (double)W > 0.6) ? 1 : 0;
        const int area = (r[4] ? (W - r[2] + r[0]) : (r[2] - r[0])) * (r[3] - r[1]);
        ow[i] = (float)area;
        if (area > max_area) max_area = area;
        for (int y = r[1]; y < r[3]; ++y)
            for (int x = 0; x < W; ++x) {
                const int in = r[4] ? (x < r[0] || x >= r[2]) : (x >= r[0] && x < r[2]);
                if (in) { mask[y * W + x] = cls[i]; mask[(size_t)H * W + y * W + x] = cdep[i]; }
            }
    }
    if (weight) {
        for (int i = 0; i < N; ++i) {
            volatile float q = ow[i] / (float)max_area;
            ow[i] = 3.f - q;
        }
        for (int y = 0; y < H; ++y)
            for (int x = 0; x < W; ++x) {
                volatile float sum = 0.f;
                for (int i = 0; i < N; ++i) {
                    const int* r = rect + i * 5;
                    const int in = y >= r[1] && y < r[3] && (r[4] ? (x < r[0] || x >= r[2]) : (x >= r[0] && x < r[2]));
                    sum = sum + (in ? ow[i] : 0.f);
                }
                weight[y * W + x] = expf(sum);
            }
    }
    free(rect); free(ow); free(cdep); free(cls);
}
