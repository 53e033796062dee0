"""ctypes wrapper of oracle/liboracle.so (C restatement of projection / points-in-boxes / voxel index)
plus NumPy restatements of LiDARUtility.  TEST ORACLE ONLY."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liboracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            subprocess.run(["make", "-C", _HERE], check=True)
        _lib = C.CDLL(_SO)
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def range_project(points, H=32, W=1024, min_depth=1.45, max_depth=80.0, fov_up=10.0, fov_down=-30.0):
    """points float32 [M,4] -> (image [H,W,6], grid int32 [M,2], winner int32 [H,W])"""
    pts = np.ascontiguousarray(points, dtype=np.float32)
    M = pts.shape[0]
    out = np.zeros((H, W, 6), np.float32)
    grid = np.zeros((M, 2), np.int32)
    win = np.zeros((H, W), np.int32)
    lib().oracle_range_project(_p(pts), M, H, W, C.c_float(min_depth), C.c_float(max_depth), C.c_float(fov_up),
                               C.c_float(fov_down), _p(out), _p(grid), _p(win))
    return out, grid, win


def range_project_f64(points, H=32, W=1024, min_depth=1.45, max_depth=80.0, fov_up=10.0, fov_down=-30.0):
    """points float64 [M,4] -> (image float32 [H,W,6], winner int32 [H,W]); common.py:26-91 on a float64 array"""
    pts = np.ascontiguousarray(points, dtype=np.float64)
    out = np.zeros((H, W, 6), np.float32)
    win = np.zeros((H, W), np.int32)
    lib().oracle_range_project_f64(_p(pts), pts.shape[0], H, W, C.c_double(min_depth), C.c_double(max_depth),
                                   C.c_double(fov_up), C.c_double(fov_down), _p(out), _p(win))
    return out, win


def boxes_to_mask(boxes, H=32, W=1024, fov_up=10.0, fov_down=-30.0):
    """convert_boxes_to_2d (common.py:99-181): boxes [N,8] float32 | float64 -> (boxes_2d f64 [N,4], mask f32 [2,H,W], weight f32 [H,W])"""
    assert boxes.dtype in (np.float32, np.float64) and boxes.shape[1] == 8
    bx = np.ascontiguousarray(boxes)
    N = bx.shape[0]
    b2 = np.zeros((N, 4), np.float64)
    mask = np.zeros((2, H, W), np.float32)
    w = np.zeros((H, W), np.float32)
    lib().oracle_boxes_to_mask(_p(bx), int(bx.dtype == np.float64), N, H, W, C.c_double(fov_up), C.c_double(fov_down), _p(b2),
                               _p(mask), _p(w))
    return b2, mask, w


def points_in_boxes(points, boxes):
    pts = np.ascontiguousarray(points, dtype=np.float32)
    bx = np.ascontiguousarray(boxes, dtype=np.float32)
    out = np.zeros((bx.shape[0], pts.shape[0]), np.int32)
    lib().oracle_points_in_boxes(_p(pts), _p(bx), bx.shape[0], pts.shape[0], _p(out))
    return out


def points_in_boxes_first(points, boxes):
    pts = np.ascontiguousarray(points, dtype=np.float32)
    bx = np.ascontiguousarray(boxes, dtype=np.float32)
    B, M, _ = pts.shape
    out = np.zeros((B, M), np.int32)
    lib().oracle_points_in_boxes_first(_p(pts), _p(bx), B, bx.shape[1], M, _p(out))
    return out


def voxel_index(points, rois, out_size):
    pts = np.ascontiguousarray(points, dtype=np.float32)
    bx = np.ascontiguousarray(rois, dtype=np.float32)
    out = np.zeros((bx.shape[0], pts.shape[0]), np.int32)
    lib().oracle_voxel_index(_p(pts), _p(bx), bx.shape[0], pts.shape[0], out_size[0], out_size[1], out_size[2], _p(out))
    return out


def depth_to_xyz(x_norm, ray_angles, min_depth=1.45, max_depth=80.0):
    """lidargen/utils/lidar.py:61-128: denormalize -> revert_depth(log_depth) -> mask -> to_xyz.  NumPy fp32."""
    nd = ((x_norm.astype(np.float32) + 1) / 2).astype(np.float32)
    metric = (np.exp2(nd * np.float32(np.log2(max_depth + 1))) - 1).astype(np.float32)
    mask = ((metric > min_depth) & (metric < max_depth)).astype(np.float32)
    metric = metric * mask
    phi, th = ray_angles[:, [0]], ray_angles[:, [1]]
    m2 = ((metric > min_depth) & (metric < max_depth)).astype(np.float32)
    xyz = np.concatenate([metric * np.cos(phi) * np.cos(th), metric * np.cos(phi) * np.sin(th), metric * np.sin(phi)], 1)
    return metric, (xyz * m2).astype(np.float32)


# ---------------------------------------------------------------------------------------------------------
# the reference's OWN roiaware_pool3d code (oracle/_ref, built by oracle/Makefile from /root/reference in the
# build container; the prebuilt .so files travel to the GPU box).  TEST INFRASTRUCTURE ONLY.
# ---------------------------------------------------------------------------------------------------------
_REF_CPU = os.path.join(_HERE, "_ref", "libref_roiaware_cpu.so")
_REF_CUDA = os.path.join(_HERE, "_ref", "libref_roiaware_cuda.so")
_ref_cpu = None
_ref_cuda = None


def ref_cpu_available() -> bool:
    return os.path.exists(_REF_CPU)


def ref_cuda_available() -> bool:
    return os.path.exists(_REF_CUDA)


def ref_points_in_boxes_cpu(points, boxes):
    """roiaware_pool3d.cpp:144-168 itself (no 0.2 m enlargement: that is the Python wrapper's job)."""
    global _ref_cpu
    if _ref_cpu is None:
        _ref_cpu = C.CDLL(_REF_CPU)
    pts = np.ascontiguousarray(points, dtype=np.float32)
    bx = np.ascontiguousarray(boxes, dtype=np.float32)
    out = np.zeros((bx.shape[0], pts.shape[0]), np.int32)
    rc = _ref_cpu.ref_points_in_boxes_cpu(_p(bx), _p(pts), bx.shape[0], pts.shape[0], _p(out))
    assert rc == 1
    return out


def ref_cuda_lib():
    """ctypes handle of the reference's CUDA kernels (device pointers in, device pointers out)."""
    global _ref_cuda
    if _ref_cuda is None:
        _ref_cuda = C.CDLL(_REF_CUDA)
    return _ref_cuda
