// oracle/ref_driver_cuda.cu -- TEST INFRASTRUCTURE ONLY.
// extern "C" doors into the reference's own CUDA kernels (lidargen/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu).
// The reference file is self-contained (math.h / stdio.h only); oracle/Makefile passes its path as REF_KERNEL_CU and it is
// #included below, unmodified and from where it lies, so its __global__ functions are visible here
// -> oracle/_ref/libref_roiaware_cuda.so.  Used by the
// `-m gpu` tests as a second pin of points_in_boxes_gpu / the voxel-index encoder (the arithmetic the reference runs on
// the device, nvcc default flags, i.e. with FMA contraction and CUDA's cosf/sinf).
#include <cuda_runtime.h>
#include REF_KERNEL_CU

extern "C" int ref_points_in_boxes_gpu(int B, int N, int M, const float* boxes_dev, const float* pts_dev, int* out_dev) {
    cudaMemset(out_dev, 0xFF, (size_t)B * M * sizeof(int));      // the Python wrapper fills -1 (roiaware_pool3d_utils.py:37)
    points_in_boxes_launcher(B, N, M, boxes_dev, pts_dev, out_dev);   // roiaware_pool3d_kernel.cu:339-355
    return (int)cudaDeviceSynchronize();
}

extern "C" int ref_voxel_index(int N, int M, int ox, int oy, int oz, const float* rois_dev, const float* pts_dev,
                               int* mask_dev) {
    cudaMemset(mask_dev, 0xFF, (size_t)N * M * sizeof(int));     // as roiaware_pool3d_launcher does (kernel.cu:214-215)
    dim3 blocks(DIVUP(M, THREADS_PER_BLOCK), N), threads(THREADS_PER_BLOCK);
    generate_pts_mask_for_box3d<<<blocks, threads>>>(N, M, ox, oy, oz, rois_dev, pts_dev, mask_dev);   // kernel.cu:39-75
    return (int)cudaDeviceSynchronize();
}
