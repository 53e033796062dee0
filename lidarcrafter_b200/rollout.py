"""GPU-resident glue of the autoregressive 4D rollout (reference: tools/vis_tools/utils/pipe_related.py:28-95,243-288 and
tools/vis_tools/utils/common.py:59-222).  Between two ``sample()`` calls the reference goes back to NumPy on the host
(ego-motion warp of the background points, pasting of the rotated object points, re-projection, foreground deletion,
3x points_in_boxes_cpu per frame); here the point sets stay on the device and use the projection / rasteriser /
points-in-boxes kernels of libb200lidar.  Point sets are RAGGED; they live in fixed-capacity buffers with a device-side
row count (``PointSet``): compaction is a cumsum + scatter, the projection kernels read the count from device memory, so a
frame of glue issues no device->host copy and no host synchronisation (the one exception is the per-object split of the
first frame, once per clip).  Trajectory -> pose arithmetic on <= 20 poses stays on the host in float64 like the reference.
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
    future_xy = np.asarray(future_xy)       # the yaws stay in the trajectory's dtype (float32 in the sampling script)
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


class PointSet:
    """[capacity, 4] (x, y, z, intensity) float32 rows, the first ``n`` (int32 device scalar, shape [1]) valid, order kept."""

    def __init__(self, buf: torch.Tensor, n: torch.Tensor):
        self.buf, self.n = buf, n

    @property
    def valid(self) -> torch.Tensor:
        return torch.arange(self.buf.shape[0], device=self.buf.device) < self.n

    def numpy(self) -> np.ndarray:          # tests / export only (synchronises)
        return self.buf[:int(self.n)].cpu().numpy()


@torch.no_grad()
def compact(rows: torch.Tensor, valid: torch.Tensor) -> PointSet:
    """order-preserving stream compaction without a host round trip: row i goes to slot cumsum(valid)[i] - 1, dropped rows to
    a dump slot behind the buffer"""
    M = rows.shape[0]
    v = valid.to(torch.int64)
    slot = torch.cumsum(v, 0) - 1
    slot = torch.where(valid, slot, torch.full_like(slot, M))
    buf = torch.zeros(M + 1, rows.shape[1], dtype=rows.dtype, device=rows.device)
    buf.index_copy_(0, slot, rows)
    return PointSet(buf[:M], v.sum().to(torch.int32).reshape(1))


@torch.no_grad()
def warp_points(points: torch.Tensor, T: np.ndarray | torch.Tensor) -> torch.Tensor:
    """pipe_related.py:244-249: homogeneous 4x4 ego-motion warp of [M,4] (x,y,z,intensity) points -> FLOAT64 [M,4] (the
    reference keeps the float64 result of ``Ts @ homo`` and re-projects THAT)."""
    Tm = torch.as_tensor(T, dtype=torch.float64, device=points.device)
    xyz1 = torch.cat([points[:, :3].double(), torch.ones_like(points[:, :1], dtype=torch.float64)], dim=1)
    out = (Tm @ xyz1.T).T.contiguous()
    out[:, 3] = points[:, 3].double()
    return out


@torch.no_grad()
def rotate_points_along_z(points: torch.Tensor, angle) -> torch.Tensor:
    """lidargen/dataset/utils.py:37-59 (points [M,3] float32): one angle, or one angle per point (tensor [M])."""
    a = torch.as_tensor(angle, dtype=torch.float32, device=points.device)
    c, s = torch.cos(a), torch.sin(a)
    x, y = points[:, 0], points[:, 1]
    return torch.stack([x * c - y * s, x * s + y * c, points[:, 2]], dim=1)


@torch.no_grad()
def warp_lidar_future(P: torch.Tensor, future_xy: np.ndarray, i: int, z0: float = 0.0) -> torch.Tensor:
    """common.py:59-113, frame i only: the first frame's background [M,4] in the LiDAR frame of future pose i (float32)."""
    future_xy = np.asarray(future_xy)
    offsets = np.vstack((future_xy[0:1], future_xy[1:] - future_xy[:-1]))
    yaw = _yaws(offsets)[i]
    c, s = np.float32(np.cos(yaw)), np.float32(np.sin(yaw))
    t = torch.tensor([future_xy[i][0], future_xy[i][1], z0], dtype=P.dtype, device=P.device)
    tr = P[:, :3] - t
    # translated.dot(R), R = [[c, -s, 0], [s, c, 0], [0, 0, 1]]
    rot = torch.stack([tr[:, 0] * float(c) + tr[:, 1] * float(s), tr[:, 0] * float(-s) + tr[:, 1] * float(c), tr[:, 2]], dim=1)
    return torch.cat([rot, P[:, 3:4]], dim=1)


def div255(x: torch.Tensor) -> torch.Tensor:
    """x / 255 as a TRUE division: with a Python-scalar divisor torch's CUDA kernel multiplies by the reciprocal (one bit off
    NumPy's quotient), and the reference's reflectance round trip `/ 255 ... * 255` is compared bit for bit"""
    return x / torch.full((), 255.0, dtype=x.dtype, device=x.device)


@torch.no_grad()
def image_points(xyz: torch.Tensor, intensity: torch.Tensor, keep: torch.Tensor | None = None):
    """xyz [3,H,W], intensity [1,H,W] (0..255 scale), keep [H,W] bool -> (rows [H*W,4], valid = ||xyz|| > 1e-2), the
    `stack(...).reshape(-1, 4)` + distance filter of pipe_related.py:70-75,273-279"""
    if keep is not None:
        k = keep[None].to(xyz.dtype)
        xyz, intensity = xyz * k, intensity * k
    rows = torch.cat([xyz, intensity], dim=0).reshape(4, -1).T.contiguous()
    return rows, torch.linalg.vector_norm(rows[:, :3], dim=1) > 1e-2


@torch.no_grad()
def refine_next_frame_points(points64: torch.Tensor, n: torch.Tensor, condition_mask: torch.Tensor, H: int = 32,
                             W: int = 1024, min_depth: float = 1.45, max_depth: float = 80.0, fov_up: float = 10.0,
                             fov_down: float = -30.0):
    """pipe_related.py:271-280: re-project the warped (float64) background, drop occluded returns and everything under the
    future boxes' 2-D masks -> (rows [H*W,4] float32, valid [H*W])"""
    img = ops.load_points_as_images(points=points64, H=H, W=W, min_depth=min_depth, max_depth=max_depth, fov_up=fov_up,
                                    fov_down=fov_down, npts=n)                                   # [H,W,6] on device
    img = (img * img[..., 5:6]).permute(2, 0, 1)
    return image_points(img[:3], div255(img[3:4]) * 255, ~(condition_mask[0] > 0))


@torch.no_grad()
def paste_objects(obj_points: torch.Tensor, obj_intensity: torch.Tensor, obj_box: torch.Tensor, boxes_3d: torch.Tensor):
    """pipe_related.py:257-267 for ALL objects at once: canonical object points [P,3] of box obj_box[p] rotated by the future
    yaw and moved to the future centre -> [P,4]"""
    b = boxes_3d.to(device=obj_points.device, dtype=torch.float32)[obj_box]
    p = rotate_points_along_z(obj_points, b[:, 6]) + b[:, :3]
    return torch.cat([p, obj_intensity[:, None]], dim=1)


@torch.no_grad()
def get_next_frame_points(background: PointSet, obj_points, obj_intensity, obj_box, fut_boxes_3d, T, condition_mask,
                          H: int = 32, W: int = 1024, min_depth: float = 1.45, max_depth: float = 80.0,
                          fov_up: float = 10.0, fov_down: float = -30.0) -> PointSet:
    """pipe_related.py:243-269: warp the background by the ego motion, re-project it (the reference round-trips through
    CustomDataset -> load_points_as_images, dropping occluded / masked returns), paste the rotated objects."""
    bg = warp_points(background.buf, T)
    rows, valid = refine_next_frame_points(bg, background.n, condition_mask, H, W, min_depth, max_depth, fov_up, fov_down)
    fg = paste_objects(obj_points, obj_intensity, obj_box, torch.as_tensor(fut_boxes_3d))
    return compact(torch.cat([rows, fg], dim=0), torch.cat([valid, torch.ones(fg.shape[0], dtype=torch.bool, device=fg.device)]))


@torch.no_grad()
def delete_fg_points(points: PointSet, boxes_3d) -> PointSet:
    """pipe_related.py:282-288: drop every point inside any (0.2 m enlarged) box."""
    boxes = torch.as_tensor(boxes_3d, dtype=torch.float32)
    if boxes.shape[0] == 0:
        return points
    m = ops.points_in_boxes_cpu(points.buf[:, :3].contiguous(), boxes[:, :7].clone())
    return compact(points.buf, (m.sum(dim=0) == 0) & points.valid)


@torch.no_grad()
def remove_ego_points(rows: torch.Tensor, center_radius: float = 2.0) -> torch.Tensor:
    """pipe_related.py:11-13 -> keep mask"""
    return ~((rows[:, 0].abs() < center_radius) & (rows[:, 1].abs() < center_radius))


@torch.no_grad()
def extract_object_points(rows: torch.Tensor, keep: torch.Tensor, boxes_3d):
    """pipe_related.py:47-68: per box, the points inside it in the box's canonical frame -> (points [P,3], intensity [P],
    box index [P]) ordered box by box, point order kept.  The ONE host synchronisation of a clip (P is data dependent)."""
    boxes = torch.as_tensor(boxes_3d, dtype=torch.float32).to(rows.device)
    m = (ops.points_in_boxes_cpu(rows[:, :3].contiguous(), boxes[:, :7].clone()) > 0) & keep[None]
    k, j = torch.nonzero(m, as_tuple=True)                  # row-major: box by box
    p = rows[j]
    canon = rotate_points_along_z(p[:, :3] - boxes[k, :3], -boxes[k, 6])
    return canon, p[:, 3].contiguous(), k
