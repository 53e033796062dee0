// oracle/ref_driver.cpp -- TEST INFRASTRUCTURE ONLY.
// extern "C" doors into the reference's own CPU code (lidargen/ops/roiaware_pool3d/src/roiaware_pool3d.cpp,
// compiled unmodified next to this file by oracle/Makefile -> oracle/_ref/libref_roiaware_cpu.so).
#include <stdio.h>
#include <stdlib.h>
#include <torch/extension.h>   // the shim (oracle/ref_shim)

int points_in_boxes_cpu(at::Tensor boxes_tensor, at::Tensor pts_tensor, at::Tensor pts_indices_tensor);   // roiaware_pool3d.cpp:144

// the reference .cpp references its CUDA launchers; this CPU-only library never calls them
#define REF_NO_CUDA(name) { fprintf(stderr, "oracle/_ref: " name " needs the CUDA build\n"); abort(); }
void roiaware_pool3d_launcher(int, int, int, int, int, int, int, const float*, const float*, const float*, int*, int*,
                              float*, int) REF_NO_CUDA("roiaware_pool3d_launcher")
void roiaware_pool3d_backward_launcher(int, int, int, int, int, int, const int*, const int*, const float*, float*, int)
    REF_NO_CUDA("roiaware_pool3d_backward_launcher")
void points_in_boxes_launcher(int, int, int, const float*, const float*, int*) REF_NO_CUDA("points_in_boxes_launcher")

extern "C" int ref_points_in_boxes_cpu(const float* boxes, const float* pts, int N, int M, int* out) {
    at::Tensor b, p, o;
    b.ptr = const_cast<float*>(boxes); b.dims[0] = N; b.dims[1] = 7;
    p.ptr = const_cast<float*>(pts);   p.dims[0] = M; p.dims[1] = 3;
    o.ptr = out;                       o.dims[0] = N; o.dims[1] = M;
    return points_in_boxes_cpu(b, p, o);
}
