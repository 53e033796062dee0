"""Box -> range-image conditioning (reference: lidargen/dataset/transforms_3d/common.py:99-216 and the batch
preprocessors of tools/vis_tools/functions/lidargen_sampler.py:35-125).

Runs once per frame on <= 13 boxes: host-side NumPy / torch, not a per-step kernel.  ``convert_boxes_to_2d`` is
restated vectorised (the reference loops over boxes in Python); integer pixel rectangles are identical to the
reference (tests/test_layout_ops.py compares against goldens of the unmodified function)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F


def convert_points_to_2d(points: np.ndarray, H: int = 64, W: int = 2048, min_depth: float = 1.45, max_depth: float = 80.0,
                         fov_up: float = 10.0, fov_down: float = -30.0) -> np.ndarray:
    """common.py:184-216 -> normalised (grid_w / W, grid_h / H) per point, same dtype flow as the reference."""
    xyz = points[:, :3]
    x, y, z = xyz[:, [0]], xyz[:, [1]], xyz[:, [2]]
    depth = np.linalg.norm(xyz, ord=2, axis=1, keepdims=True) + 1e-6
    h_up, h_down = np.deg2rad(fov_up), np.deg2rad(fov_down)
    elevation = np.arcsin(z / depth) + abs(h_down)
    grid_h = np.floor((1 - elevation / (h_up - h_down)) * H).clip(0, H - 1) / H
    azimuth = -np.arctan2(y, x)
    grid_w = np.floor(((azimuth / np.pi + 1) / 2 % 1) * W).clip(0, W - 1) / W
    return np.concatenate((grid_w, grid_h), axis=1)


def convert_boxes_to_2d(boxes_3d: np.ndarray, H: int = 64, W: int = 2048, min_depth: float = 1.45, max_depth: float = 80.0,
                        fov_up: float = 10.0, fov_down: float = -30.0):
    """common.py:99-181: boxes [N, >=8] (x,y,z,l,w,h,yaw,class) -> (boxes_2d [N,4], condition_mask [2,H,W],
    scene_loss_weight_map [H,W])."""
    n = boxes_3d.shape[0]
    l, w, h = boxes_3d[:, 3], boxes_3d[:, 4], boxes_3d[:, 5]
    sx = np.array([1, 1, -1, -1, 1, 1, -1, -1]) * 0.5
    sy = np.array([1, -1, -1, 1, 1, -1, -1, 1]) * 0.5
    sz = np.array([1, 1, 1, 1, -1, -1, -1, -1]) * 0.5
    local = np.stack([l[:, None] * sx, w[:, None] * sy, h[:, None] * sz], axis=1)            # [N,3,8]
    c, s = np.cos(boxes_3d[:, 6]), np.sin(boxes_3d[:, 6])
    rot = np.zeros((n, 3, 3)); rot[:, 0, 0] = c; rot[:, 0, 1] = -s; rot[:, 1, 0] = s; rot[:, 1, 1] = c; rot[:, 2, 2] = 1
    centre = boxes_3d[:, :3][:, :, None]
    corners = (rot @ local + centre).transpose(0, 2, 1).reshape(-1, 3)
    c_depth = np.linalg.norm(centre, ord=2, axis=1, keepdims=True) + 1e-6
    uv = convert_points_to_2d(corners, H, W, min_depth, max_depth, fov_up, fov_down).reshape(n, 8, 2)
    boxes_2d = np.stack([uv[..., 0].min(1), uv[..., 1].min(1), uv[..., 0].max(1), uv[..., 1].max(1)], axis=1)
    mask = np.zeros([2, H, W], dtype=np.float32)
    weight = np.zeros([H, W, n], dtype=np.float32)
    areas = []
    for i, (x1, y1, x2, y2) in enumerate(boxes_2d):
        x1, x2, y1, y2 = int(x1 * W), int(x2 * W), int(y1 * H), int(y2 * H)
        if (x2 - x1) / W > 0.6:          # box straddles the azimuth seam: fill both ends (common.py:152-163)
            cols = [slice(0, x1), slice(x2, W)]
            areas.append((W - x2 + x1) * (y2 - y1))
        else:
            cols = [slice(x1, x2)]
            areas.append((x2 - x1) * (y2 - y1))
        for cs in cols:
            mask[0, y1:y2, cs] = boxes_3d[i, 7]
            mask[1, y1:y2, cs] = c_depth[i, 0, 0]
            weight[y1:y2, cs, i] = 1.0
    areas = np.array(areas, dtype=np.float32)
    weight = weight * (3 - areas / np.max(areas))[None, None, :]
    return boxes_2d, mask, np.exp(weight.sum(-1))


def preprocess_condition_mask(condition_mask: torch.Tensor, lidar_utils, num_classes: int = 9) -> torch.Tensor:
    """lidargen_sampler.py:70-81: [B,2,H,W] (class id, centre depth) -> concat_cond [B, num_classes+1, H, W]."""
    one_hot = F.one_hot(condition_mask[:, 0].long(), num_classes=num_classes).permute(0, 3, 1, 2).float()
    depth = lidar_utils.convert_depth(condition_mask[:, 1].unsqueeze(1))
    return torch.cat([one_hot, depth], dim=1)


def preprocess_autoregressive_cond(ar: torch.Tensor, lidar_utils, resolution, with_reflectance: bool = False):
    """lidargen_sampler.py:83-99 ('nuscenes-auto-reg-v2': depth only) -> [-1,1] normalised [B,1|2,H,W]."""
    x = [lidar_utils.convert_depth(ar[:, 0].unsqueeze(1))]
    if with_reflectance:
        x.append(ar[:, 1].unsqueeze(1))
    x = lidar_utils.normalize(torch.cat(x, dim=1))
    return F.interpolate(x, size=tuple(resolution), mode="nearest-exact")
