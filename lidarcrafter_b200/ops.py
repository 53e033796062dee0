"""Point-cloud ops of the hot path on the B200 (reference: lidargen/dataset/transforms_3d/common.py:26-91 and
lidargen/ops/roiaware_pool3d/roiaware_pool3d_utils.py:9-41).  Same function names / argument meaning as the
reference; tensors stay on the device (numpy inputs are uploaded, numpy outputs returned for numpy inputs)."""
from __future__ import annotations

import numpy as np
import torch

from . import _lib


def _stream():
    return 0 if _lib.compute_device() == "cpu" else torch.cuda.current_stream().cuda_stream


def _dev(x, dtype=torch.float32):
    is_np = isinstance(x, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(x)) if is_np else x
    return t.to(device=_lib.compute_device(), dtype=dtype).contiguous(), is_np


def load_points_as_images(point_path: str = None, points=None, scan_unfolding: bool = False, H: int = 64, W: int = 2048,
                          min_depth: float = 1.45, max_depth: float = 80.0, fov_up: float = 10.0,
                          fov_down: float = -30.0, custom_feat_dim: int = 0, return_grid: bool = False, npts=None):
    """common.py:26-91 -> float32 [H, W, 6] (x, y, z, intensity, depth, mask); nearest return per pixel.
    Batched form: points [F, M, 4] -> [F, H, W, 6]; ``npts`` (int32 device tensor [F]) = valid rows per frame (ragged
    clouds in a fixed-capacity buffer, no host round trip).  float64 points take the reference's float64 dtype flow
    (b200_range_project_f64: what the temporal glue feeds the function), float32 / anything else the float32 one."""
    if scan_unfolding:
        raise NotImplementedError("scan_unfolding=True is not used by any nuScenes config (configs/*: False)")
    if custom_feat_dim:
        raise NotImplementedError("custom_feat_dim != 0 is not on the hot path")
    assert point_path is not None or points is not None, "Either point_path or points must be provided."
    if point_path is not None:
        points = np.fromfile(point_path, dtype=np.float32).reshape(-1, 5)[:, :4]
    is64 = (points.dtype == np.float64) if isinstance(points, np.ndarray) else (points.dtype == torch.float64)
    pts, is_np = _dev(points, torch.float64 if is64 else torch.float32)
    single = pts.dim() == 2
    if single:
        pts = pts[None]
    assert pts.shape[-1] == 4
    F, M, _ = pts.shape
    out = torch.empty(F, H, W, 6, device=pts.device)
    zbuf = torch.empty(F, H, W, dtype=torch.int64, device=pts.device)
    grid = torch.empty(F, M, 2, dtype=torch.int32, device=pts.device) if return_grid else None
    n_ptr = 0 if npts is None else npts.to(device=pts.device, dtype=torch.int32).contiguous().data_ptr()
    if is64:
        if return_grid:
            raise NotImplementedError("return_grid with float64 points")
        win = torch.empty(F, H, W, dtype=torch.int32, device=pts.device)
        _lib.get_lib().range_project_f64(pts.data_ptr(), n_ptr, out.data_ptr(), zbuf.data_ptr(), win.data_ptr(), F, M, H, W,
                                         float(min_depth), float(max_depth), float(fov_up), float(fov_down), _stream())
    else:
        _lib.get_lib().range_project(pts.data_ptr(), n_ptr, out.data_ptr(), 0 if grid is None else grid.data_ptr(),
                                     zbuf.data_ptr(), F, M, H, W, float(min_depth), float(max_depth), float(fov_up),
                                     float(fov_down), _stream())
    if single:
        out = out[0]
        grid = None if grid is None else grid[0]
    if is_np:
        out = out.cpu().numpy()
        grid = None if grid is None else grid.cpu().numpy()
    return (out, grid) if return_grid else out


def points_in_boxes_cpu(points, boxes):
    """roiaware_pool3d_utils.py:9-25 (name kept; runs on the GPU): boxes are enlarged by 0.2 m, MARGIN 1e-2.
    points (M,3), boxes (N,7) -> int32 (N, M) of 0/1."""
    assert boxes.shape[1] == 7 and points.shape[1] == 3
    pts, is_np = _dev(points)
    bx, _ = _dev(boxes)
    bx = bx.clone()
    bx[:, 3:6] += 0.2
    if not isinstance(boxes, np.ndarray) or boxes.dtype == np.float32:
        # the reference enlarges the caller's boxes IN PLACE: torch tensors always, NumPy arrays when they are float32
        # (check_numpy_to_torch: torch.from_numpy(x).float() shares memory with a float32 array, dataset/utils.py:13-16)
        boxes[:, 3:6] += 0.2
    out = torch.zeros(bx.shape[0], pts.shape[0], dtype=torch.int32, device=pts.device)
    if bx.shape[0] and pts.shape[0]:
        _lib.get_lib().points_in_boxes(pts.data_ptr(), bx.data_ptr(), out.data_ptr(), bx.shape[0], pts.shape[0],
                                       _stream())
    return out.cpu().numpy() if is_np else out


def points_in_boxes_gpu(points, boxes):
    """roiaware_pool3d_utils.py:28-41: points (B,M,3), boxes (B,T,7) -> int32 (B,M), first containing box or -1."""
    assert boxes.shape[0] == points.shape[0] and boxes.shape[2] == 7 and points.shape[2] == 3
    pts, _ = _dev(points)
    bx, _ = _dev(boxes)
    B, M, _ = pts.shape
    out = torch.full((B, M), -1, dtype=torch.int32, device=pts.device)
    _lib.get_lib().points_in_boxes_first(pts.data_ptr(), bx.data_ptr(), out.data_ptr(), B, bx.shape[1], M,
                                         _stream())
    return out


def voxel_index(points, rois, out_size):
    """generate_pts_mask_for_box3d (roiaware_pool3d_kernel.cu:39-75): (M,3), (N,7) -> int32 (N,M): -1 or x<<16|y<<8|z."""
    pts, is_np = _dev(points)
    bx, _ = _dev(rois)
    ox, oy, oz = (out_size,) * 3 if isinstance(out_size, int) else out_size
    out = torch.empty(bx.shape[0], pts.shape[0], dtype=torch.int32, device=pts.device)
    _lib.get_lib().voxel_index(pts.data_ptr(), bx.data_ptr(), out.data_ptr(), bx.shape[0], pts.shape[0], ox, oy, oz,
                               _stream())
    return out.cpu().numpy() if is_np else out
