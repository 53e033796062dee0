"""Box -> range-image conditioning (reference: lidargen/dataset/transforms_3d/common.py:99-216 and the batch
preprocessors of tools/vis_tools/functions/lidargen_sampler.py:35-125).

``convert_boxes_to_2d`` runs on the device (b200_boxes_to_mask: all frames / samples of a batch in one launch pair);
the <= 13 rows of box scaling / padding around it are host NumPy in the reference's own dtypes.  Pixel rectangles, masks
and 2-D boxes are bit-identical to the reference (tests/golden/boxes2d.npz, temporal.npz: goldens of the unmodified
functions for float32 and float64 boxes)."""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

from . import _lib


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


def convert_boxes_to_2d(boxes_3d, H: int = 64, W: int = 2048, min_depth: float = 1.45, max_depth: float = 80.0,
                        fov_up: float = 10.0, fov_down: float = -30.0):
    """common.py:99-181: boxes [N, 8] (x,y,z,l,w,h,yaw,class), float32 or float64 -> (boxes_2d [N,4] float64,
    condition_mask [2,H,W] float32, scene_loss_weight_map [H,W] float32) on the device (b200_boxes_to_mask: corner
    projection in the reference's dtype flow + one rasteriser pass).  NumPy in -> NumPy out, torch in -> device tensors;
    batched form [F,N,8] -> ([F,N,4], [F,2,H,W], [F,H,W])."""
    is_np = isinstance(boxes_3d, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(boxes_3d)) if is_np else boxes_3d
    if t.dtype not in (torch.float32, torch.float64):
        t = t.double()
    single = t.dim() == 2
    t = (t[None] if single else t)[..., :8].to(_lib.compute_device()).contiguous()
    F_, N = t.shape[:2]
    lib = _lib.get_lib()
    dev = t.device
    b2 = torch.empty(F_, N, 4, dtype=torch.float64, device=dev)
    mask = torch.empty(F_, 2, H, W, dtype=torch.float32, device=dev)
    weight = torch.empty(F_, H, W, dtype=torch.float32, device=dev)
    ws = torch.empty(max(int(lib.boxes_to_mask_workspace(F_, N)), 8), dtype=torch.uint8, device=dev)
    lib.boxes_to_mask(t.data_ptr(), 1 if t.dtype == torch.float64 else 0, F_, N, H, W, float(fov_up), float(fov_down),
                      b2.data_ptr(), mask.data_ptr(), weight.data_ptr(), ws.data_ptr(), _lib.current_stream(dev))
    if single:
        b2, mask, weight = b2[0], mask[0], weight[0]
    if is_np:
        return b2.cpu().numpy(), mask.cpu().numpy(), weight.cpu().numpy()
    return b2, mask, weight


# ---- NuscDataset.pre_process (dataset/nuscenes_dataset.py:145-236,375-421): per-frame box conditioning, <= 13 rows of host
# float arithmetic in the reference's own NumPy dtypes + the device rasteriser above ----
CLASS_NAMES = ("car", "truck", "construction_vehicle", "bus", "trailer", "motorcycle", "bicycle", "pedestrian")
POINTS_RANGE = (-80, -80, -8, 80, 80, 8)


def scale_boxes_3d(boxes_3d: np.ndarray, points_range=POINTS_RANGE) -> np.ndarray:
    """nuscenes_dataset.py:145-158: [N,7+] (x,y,z,l,w,h,yaw,extra...) -> float64 [N,8+]: centre over the half extent of the point
    range, log sizes, (sin, cos) of the yaw, extra columns.  The divisions / log / sin / cos run in the INPUT's dtype (as in
    the reference, which edits its argument in place) and are widened to float64 only when stored."""
    src = np.asarray(boxes_3d)
    n, d = src.shape
    out = np.zeros((n, d + 1))
    for axis, lo in enumerate(points_range[:3]):            # Python-int divisors keep a float32 input in float32
        out[:, axis] = src[:, axis] / (0 - lo)
    out[:, 3:6] = np.log(src[:, 3:6] + 1e-6)
    out[:, 6], out[:, 7] = np.sin(src[:, 6]), np.cos(src[:, 6])
    out[:, 8:] = src[:, 7:]
    return out


def encoding_boxes_3d(box: np.ndarray, unique_mode: bool = True, points_range=POINTS_RANGE) -> np.ndarray:
    """nuscenes_dataset.py:194-214 for ONE box (x,y,z,w,h,l,yaw) -> float32 [6] (unique_mode) | [8]:
    (planar distance of the normalised centre, normalised z, log sizes, then the viewing-relative yaw | azimuth bin + sin / cos yaw)"""
    x, y, z, yaw = box[0], box[1], box[2], box[6]
    lo = points_range[:3]
    enc = np.zeros(8, dtype=np.float32)
    enc[0] = np.linalg.norm(np.array([x / (0 - lo[0]), y / (0 - lo[1])]), ord=2, axis=0, keepdims=True)[0]
    enc[1] = z / (0 - lo[2])
    enc[2:5] = np.log(np.array([box[3], box[4], box[5]]) + 1e-6)
    bearing = np.arctan2(y, x)
    if unique_mode:
        enc[5] = yaw - bearing
        return enc[:6]
    enc[5:8] = (-bearing / np.pi + 1) / 2 % 1, np.sin(yaw), np.cos(yaw)
    return enc


def _fit_rows(a: np.ndarray, rows: int) -> np.ndarray:
    """first `rows` rows of `a`, zero-padded to `rows` (float64 when padding, like the reference's np.zeros buffers)"""
    if a.shape[0] >= rows:
        return a[:rows]
    out = np.zeros((rows, a.shape[-1]))
    out[:a.shape[0]] = a
    return out


def allign_box_num(bbox_3d, bbox_2d, fg_encoding_box, expet_box_num: int = 13):
    """nuscenes_dataset.py:174-192 (name as in the reference): the layout encoder takes exactly 13 object rows -> the three
    per-object arrays cut / zero-padded to 13 rows + the validity flags"""
    valid = (np.arange(expet_box_num) < bbox_3d.shape[0]).astype(np.float64)
    return (_fit_rows(bbox_3d, expet_box_num), _fit_rows(bbox_2d, expet_box_num), _fit_rows(fg_encoding_box, expet_box_num), valid)


def class_ids(gt_names, class_names=CLASS_NAMES) -> np.ndarray:
    names = ["ego"] + list(class_names)
    return np.array([names.index(n) for n in gt_names], dtype=np.int32)


def layout_item(gt_boxes: np.ndarray, gt_names, H: int = 32, W: int = 1024, min_depth: float = 1.45, max_depth: float = 80.0,
                fov_up: float = 10.0, fov_down: float = -30.0, class_names=CLASS_NAMES, host: bool = True) -> dict:
    """pre_process for the tasks 'layout_cond' / 'autoregressive_generation' (nuscenes_dataset.py:375-421): gt_boxes [N+1,7]
    with the ego row first + names -> the conditioning entries of one dataset item.  ``host=False`` keeps the rasteriser
    outputs (gt_boxes_2d, condition_mask, scene_loss_weight_map) on the device as torch tensors."""
    fg = np.stack([encoding_boxes_3d(b[:7], unique_mode=False) for b in gt_boxes[1:]], axis=0)
    boxes8 = np.concatenate((gt_boxes, class_ids(gt_names, class_names).reshape(-1, 1).astype(np.float32)), axis=1)
    if host:
        b2, mask, weight = convert_boxes_to_2d(boxes8, H, W, min_depth, max_depth, fov_up, fov_down)
        scaled, b2p, fgp, valid = allign_box_num(scale_boxes_3d(boxes8.copy())[1:], b2[1:], fg)
    else:
        b2, mask, weight = convert_boxes_to_2d(torch.from_numpy(boxes8), H, W, min_depth, max_depth, fov_up, fov_down)
        scaled, _, fgp, valid = allign_box_num(scale_boxes_3d(boxes8.copy())[1:], np.zeros((boxes8.shape[0] - 1, 4)), fg)
        b2p = torch.zeros(13, 4, dtype=torch.float64, device=b2.device)
        n = min(13, b2.shape[0] - 1)
        b2p[:n] = b2[1:1 + n]
    return dict(gt_boxes=boxes8, scaled_gt_boxes=scaled, fg_encoding_box=fgp, gt_boxes_2d=b2p, is_valid_obj=valid,
                condition_mask=mask, scene_loss_weight_map=weight)


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
