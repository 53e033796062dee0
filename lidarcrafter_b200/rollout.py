"""GPU-resident glue of the autoregressive 4D rollout (reference: tools/vis_tools/utils/pipe_related.py:243-288 and
tools/vis_tools/utils/common.py:59-222).  Between two ``sample()`` calls the reference goes back to NumPy on the host
(ego-motion warp of the background points, pasting of the rotated object points, re-projection, foreground deletion,
3x points_in_boxes_cpu per frame); here the point sets stay on the device and use the projection / points-in-boxes
kernels of libb200lidar.  Trajectory -> pose arithmetic on <= 20 poses stays on the host in float64 like the reference.
"""
from __future__ import annotations

import numpy as np
import torch

from . import ops


def _yaws(offsets: np.ndarray) -> np.ndarray:
    yaw = np.arctan2(offsets[:, 1], offsets[:, 0]) - np.pi / 2
    yaw[np.linalg.norm(offsets, axis=1) < 1e-1] = 0.0
    return yaw


def compute_inter_frame_transforms(future_xy: np.ndarray, z0: float = 0.0) -> np.ndarray:
    """common.py:172-222: (T,2) ego trajectory in the first LiDAR frame -> (T,4,4) maps frame i -> frame i+1."""
    future_xy = np.asarray(future_xy, dtype=np.float64)
    offsets = np.vstack((future_xy[0:1], future_xy[1:] - future_xy[:-1]))
    yaws = _yaws(offsets)
    poses = [np.eye(4)]
    for (x, y), yaw in zip(future_xy, yaws):
        P = np.eye(4)
        c, s = np.cos(yaw), np.sin(yaw)
        P[:3, :3] = [[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]]
        P[:3, 3] = [x, y, z0]
        poses.append(P)
    return np.stack([np.linalg.inv(poses[i + 1]) @ poses[i] for i in range(len(future_xy))])


def warp_boxes_future(boxes0: np.ndarray, traj_obj: np.ndarray, traj_ego: np.ndarray, z_e: float = 0.0) -> np.ndarray:
    """common.py:116-169: (K,7) boxes + per-object / ego (dx,dy) trajectories -> (K,N,7) boxes in each future LiDAR frame."""
    K, N = traj_obj.shape[:2]
    out = np.zeros((K, N, 7), dtype=boxes0.dtype)
    yaw_ego = _yaws(np.vstack((traj_ego[0:1], traj_ego[1:] - traj_ego[:-1])))
    for k in range(K):
        x0, y0, z0, w, h, l, yaw0 = boxes0[k]
        step = traj_obj[k, 1:] - traj_obj[k, :-1]
        heading = np.arctan2(step[:, 1], step[:, 0])
        still = np.linalg.norm(step, axis=1) < 1e-3
        yaw_obj = np.empty(N, dtype=boxes0.dtype)
        yaw_obj[0] = yaw0
        for i in range(1, N):
            yaw_obj[i] = yaw_obj[i - 1] if still[i - 1] else heading[i - 1]
        for i in range(N):
            d = np.array([x0 + traj_obj[k, i, 0] - traj_ego[i, 0], y0 + traj_obj[k, i, 1] - traj_ego[i, 1], z0 - z_e],
                         dtype=boxes0.dtype)
            c, s = np.cos(yaw_ego[i]), np.sin(yaw_ego[i])
            out[k, i, :3] = [c * d[0] + s * d[1], -s * d[0] + c * d[1], d[2]]
            out[k, i, 3:6] = [w, h, l]
            out[k, i, 6] = yaw_obj[i] - yaw_ego[i]
    return out


@torch.no_grad()
def warp_points(points: torch.Tensor, T: np.ndarray | torch.Tensor) -> torch.Tensor:
    """pipe_related.py:244-249: homogeneous 4x4 ego-motion warp of [M,4] (x,y,z,intensity) points (fp64 math)."""
    Tm = torch.as_tensor(T, dtype=torch.float64, device=points.device)
    xyz1 = torch.cat([points[:, :3].double(), torch.ones_like(points[:, :1], dtype=torch.float64)], dim=1)
    out = (Tm @ xyz1.T).T
    out[:, 3] = points[:, 3].double()
    return out.to(points.dtype)


@torch.no_grad()
def rotate_points_along_z(points: torch.Tensor, angle: float) -> torch.Tensor:
    """lidargen/dataset/utils.py rotate_points_along_z for one box (points [M,3])."""
    c, s = float(np.cos(angle)), float(np.sin(angle))
    R = torch.tensor([[c, s, 0.0], [-s, c, 0.0], [0.0, 0.0, 1.0]], dtype=points.dtype, device=points.device)
    return points @ R


@torch.no_grad()
def delete_fg_points(points: torch.Tensor, boxes_3d: torch.Tensor) -> torch.Tensor:
    """pipe_related.py:282-288: drop every point inside any (0.2 m enlarged) box."""
    if boxes_3d.shape[0] == 0:
        return points
    m = ops.points_in_boxes_cpu(points[:, :3].contiguous(), boxes_3d[:, :7].clone())
    return points[m.sum(dim=0) == 0]


@torch.no_grad()
def extract_object_points(points: torch.Tensor, boxes_3d: torch.Tensor):
    """pipe_related.py:54-68: per box, the points inside it in the box's canonical frame (+ their intensity)."""
    m = ops.points_in_boxes_cpu(points[:, :3].contiguous(), boxes_3d[:, :7].clone())
    objs, inten = [], []
    for k in range(boxes_3d.shape[0]):
        p = points[m[k] > 0]
        inten.append(p[:, 3])
        objs.append(rotate_points_along_z(p[:, :3] - boxes_3d[k, :3].to(p.dtype), -float(boxes_3d[k, 6])))
    return objs, inten


@torch.no_grad()
def get_next_frame_points(background: torch.Tensor, obj_points, obj_intensity, fut_boxes_3d: torch.Tensor, T,
                          H: int = 32, W: int = 1024, min_depth: float = 1.45, max_depth: float = 80.0,
                          fov_up: float = 10.0, fov_down: float = -30.0, condition_mask: torch.Tensor | None = None):
    """pipe_related.py:243-280: warp the background by the ego motion, re-project it (the reference round-trips through
    CustomDataset -> load_points_as_images, dropping occluded / masked returns), paste the rotated objects."""
    bg = warp_points(background, T)
    img = ops.load_points_as_images(points=bg[:, :4].float().contiguous(), H=H, W=W, min_depth=min_depth,
                                    max_depth=max_depth, fov_up=fov_up, fov_down=fov_down)     # [H,W,6] on device
    img = img * img[..., 5:6]
    if condition_mask is not None:                       # remove background under the future boxes' 2-D masks
        img = img * (~(condition_mask[0] > 0))[..., None].to(img.dtype)
    pts = img[..., :4].reshape(-1, 4)
    pts = pts[pts[:, :3].norm(dim=1) > 1e-2]
    fg = []
    for k in range(fut_boxes_3d.shape[0]):
        p = rotate_points_along_z(obj_points[k], float(fut_boxes_3d[k, 6])) + fut_boxes_3d[k, :3].to(obj_points[k].dtype)
        fg.append(torch.cat([p, obj_intensity[k][:, None]], dim=1))
    return torch.cat([pts] + fg, dim=0) if fg else pts
