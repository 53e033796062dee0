"""Evaluation-side point-cloud ops on the B200: same function names / argument meaning as
``lidargen/metrics/metric_utils.py`` (pcd2range :65-121, range2xyz :124-154, ravel_hash :28-41, sparse_quantize :44-62,
pcd2voxel_full :170-199, pcd2bev_sum :231-256, pcd2bev_bin :259-283, bev_sample :286-306, preprocess_pcd/_range :309-322).

Inputs may be NumPy arrays (uploaded; NumPy returned, like the reference) or CUDA tensors (results stay on the device).
Point clouds are float32, the dtype every caller of the reference produces (``np.fromfile(..., np.float32)``,
``np.loadtxt(..., dtype=np.float32)``); the binning arithmetic is the fp32 arithmetic NumPy >= 2 applies to such input.
There is no CPU fallback: every op runs a libb200lidar kernel.
"""
from __future__ import annotations

import ctypes as C
import math
from itertools import repeat

import numpy as np
import torch

from . import _lib

# lidargen/metrics/__init__.py:28-36 (eval settings, "do not modify")
VOXEL_SIZE = 0.05
DATA_CONFIG = {'64': {'x': [-50, 50], 'y': [-50, 50], 'z': [-3, 1]},
               '32': {'x': [-30, 30], 'y': [-30, 30], 'z': [-3, 6]}}
DATASET_CONFIG = {'kitti': {'size': [64, 1024], 'fov': [3, -25], 'depth_range': [1.0, 56.0], 'depth_scale': 6},
                  'nuscenes': {'size': [32, 1024], 'fov': [10, -30], 'depth_range': [1.0, 45.0]}}


def _stream():
    return 0 if _lib.compute_device() == "cpu" else torch.cuda.current_stream().cuda_stream


def _dev(x, dtype=torch.float32):
    is_np = isinstance(x, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(x)) if is_np else x
    return t.to(device=_lib.compute_device(), dtype=dtype).contiguous(), is_np


def _ret(t, is_np):
    return t.cpu().numpy() if is_np else t


# ---------------------------------------------------------------------------------------------------------
# projection
# ---------------------------------------------------------------------------------------------------------
def pcd2range(pcd, size, fov, depth_range, remission=None, labels=None, **kwargs):
    """metric_utils.py:65-121 -> (proj_range [H,W] fp32 with -1 fill, proj_feature or None).
    Batched form: pcd [F,M,3] (+ remission/labels [F,M]) -> [F,H,W]."""
    pts, is_np = _dev(pcd)
    single = pts.dim() == 2
    if single:
        pts = pts[None]
    if pts.shape[-1] != 3:
        # the reference takes the depth as the norm over ALL columns (np.linalg.norm(pcd, 2, axis=1), metric_utils.py:72):
        # an [M,4] cloud with intensity would silently change every range value -- pass xyz only
        raise ValueError(f"pcd2range expects xyz columns only, got {pts.shape[-1]} columns")
    pts = pts.contiguous()
    F, M, _ = pts.shape
    H, W = int(size[0]), int(size[1])
    feat, fill = None, 0.0
    if remission is not None:
        feat, fill = _dev(remission)[0].reshape(F, M), -1.0
    elif labels is not None:
        feat, fill = _dev(labels)[0].reshape(F, M), 0.0
    rng = torch.empty(F, H, W, device=pts.device)
    pf = torch.empty(F, H, W, device=pts.device) if feat is not None else None
    zbuf = torch.empty(F, H, W, dtype=torch.int64, device=pts.device)
    _lib.get_lib().pcd2range(pts.data_ptr(), 0, 0 if feat is None else feat.data_ptr(), rng.data_ptr(),
                             0 if pf is None else pf.data_ptr(), zbuf.data_ptr(), F, M, H, W, float(fov[0]), float(fov[1]),
                             float(depth_range[0]), float(depth_range[1]), float(fill), _stream())
    if single:
        rng, pf = rng[0], (None if pf is None else pf[0])
    return _ret(rng, is_np), (None if pf is None else _ret(pf, is_np))


def range2xyz(range_img, fov, depth_range, depth_scale=None, log_scale=True, **kwargs):
    """metric_utils.py:124-154: [H,W] (or [F,H,W]) fp32 range image -> float64 [3,H,W] (or [F,3,H,W])."""
    img, is_np = _dev(range_img)
    single = img.dim() == 2
    if single:
        img = img[None]
    F, H, W = img.shape
    xyz = torch.empty(F, 3, H, W, dtype=torch.float64, device=img.device)
    _lib.get_lib().range2xyz(img.data_ptr(), xyz.data_ptr(), F, H, W, float(fov[0]), float(fov[1]), float(depth_range[0]),
                             float(depth_range[1]), float(depth_scale if log_scale else 0.0), 1 if log_scale else 0, _stream())
    return _ret(xyz[0] if single else xyz, is_np)


def preprocess_pcd(pcd, **kwargs):
    """metric_utils.py:309-313"""
    pts, is_np = _dev(pcd)
    depth = torch.linalg.vector_norm(pts, dim=1)          # over ALL columns, like the reference (metric_utils.py:310)
    out = pts[(depth > kwargs['depth_range'][0]) & (depth < kwargs['depth_range'][1])]
    return _ret(out, is_np)


def preprocess_range(pcd, **kwargs):
    """metric_utils.py:316-322 -> float64 [4,H,W] = [depth image ; xyz image]"""
    pts, is_np = _dev(pcd)
    depth_img = pcd2range(pts, **kwargs)[0]
    xyz_img = range2xyz(depth_img, log_scale=False, **{k: v for k, v in kwargs.items() if k != 'log_scale'})
    img = torch.cat([depth_img[None].double(), xyz_img], dim=0)
    return _ret(img, is_np)


# ---------------------------------------------------------------------------------------------------------
# voxel quantiser
# ---------------------------------------------------------------------------------------------------------
def _quantize(coords: torch.Tensor, voxel_size, div_f32: bool):
    """-> (voxel int32 [M,D] on device, minmax as a host int32[6])"""
    M, D = coords.shape
    assert D in (2, 3), "support 2D and 3D coordinates only"
    assert coords.dtype in (torch.float32, torch.float64)
    vs = list(voxel_size) + [1.0] * (3 - D)
    voxel = torch.empty(M, D, dtype=torch.int32, device=coords.device)
    mm = torch.empty(6, dtype=torch.int32, device=coords.device)
    _lib.get_lib().quantize_coords(coords.data_ptr(), 1 if coords.dtype == torch.float64 else 0, M, D, coords.stride(0),
                                   float(vs[0]), float(vs[1]), float(vs[2]), 1 if div_f32 else 0, voxel.data_ptr(),
                                   mm.data_ptr(), _stream())
    return voxel, mm.cpu().numpy().astype(np.int32)        # the extents size the workspace: one small D2H sync


def _unique(voxel: torch.Tensor, minmax: np.ndarray, want_inverse: bool = True):
    """np.unique(ravel_hash(voxel), return_index=True, return_inverse=True) -> (coords[U,D] int32, indices, inverse)"""
    lib = _lib.get_lib()
    M, D = voxel.shape
    mm = (C.c_int32 * 6)(*[int(v) for v in minmax])
    need = lib.sparse_quantize_workspace(mm, D, M)
    if need == 0:
        raise _lib.B200LidarError("sparse_quantize: the bounding voxel grid of this cloud exceeds the 2^35-cell bitmap limit")
    ws = torch.empty(need, dtype=torch.uint8, device=voxel.device)
    uniq = torch.empty(M, D, dtype=torch.int32, device=voxel.device)
    idx = torch.empty(M, dtype=torch.int64, device=voxel.device)
    inv = torch.empty(M, dtype=torch.int64, device=voxel.device) if want_inverse else None
    n = torch.zeros(1, dtype=torch.int32, device=voxel.device)
    lib.sparse_quantize(voxel.data_ptr(), M, D, mm, ws.data_ptr(), need, uniq.data_ptr(), idx.data_ptr(),
                        0 if inv is None else inv.data_ptr(), n.data_ptr(), _stream())
    U = int(n.item())
    return uniq[:U], idx[:U], inv


def ravel_hash(x):
    """metric_utils.py:28-41: integer coordinates [M,D] -> uint64 keys (returned as int64 bit patterns on the device)."""
    t, is_np = _dev(x, dtype=torch.int32)
    assert t.dim() == 2, t.shape
    mn, mx = t.min(dim=0).values, t.max(dim=0).values
    D = t.shape[1]
    mm = np.zeros(6, np.int32)
    mm[:D], mm[3:3 + D] = mn.cpu().numpy(), mx.cpu().numpy()
    out = torch.empty(t.shape[0], dtype=torch.int64, device=t.device)
    _lib.get_lib().ravel_hash(t.data_ptr(), t.shape[0], D, (C.c_int32 * 6)(*[int(v) for v in mm]), out.data_ptr(), _stream())
    return out.cpu().numpy().view(np.uint64) if is_np else out


def sparse_quantize(coords, voxel_size=1, *, return_index: bool = False, return_inverse: bool = False):
    """metric_utils.py:44-62 (torchsparse semantics): unique voxel coordinates in ravel-hash order."""
    is_f64 = coords.dtype in (np.float64, torch.float64)      # fp64 coordinates are quantised in fp64, everything else in fp32
    c, is_np = _dev(coords, dtype=torch.float64 if is_f64 else torch.float32)
    if isinstance(voxel_size, (float, int)):
        voxel_size = tuple(repeat(voxel_size, c.shape[1]))
    assert isinstance(voxel_size, tuple) and len(voxel_size) in [2, 3]
    voxel, mm = _quantize(c, voxel_size, div_f32=False)
    uniq, idx, inv = _unique(voxel, mm, want_inverse=return_inverse)
    outputs = [_ret(uniq, is_np)]
    if return_index:
        outputs += [_ret(idx, is_np)]
    if return_inverse:
        outputs += [_ret(inv, is_np)]
    return outputs[0] if len(outputs) == 1 else outputs


def _bounds(data_type, voxel_size, dims=2):
    cfg = DATA_CONFIG[data_type]
    rng = [cfg['x'], cfg['y'], cfg['z']][:dims]
    shape = tuple(math.ceil((r[1] - r[0]) / voxel_size) for r in rng)
    minb = tuple(math.ceil(r[0] / voxel_size) for r in rng)
    return rng, shape, minb


def _bev_filter(pts: torch.Tensor, rng):
    m = (pts[:, 0] > rng[0][0]) & (pts[:, 0] < rng[0][1]) & (pts[:, 1] > rng[1][0]) & (pts[:, 1] < rng[1][1])
    return pts[m][:, :2].contiguous()


def pcd2bev_sum(data_type, *args, voxel_size=VOXEL_SIZE):
    """metric_utils.py:231-256: per data set (a list of clouds) the [X,Y] count of clouds occupying each BEV cell.
    One launch per data set: the clouds are concatenated (CSR offsets), one bitmap per cloud de-duplicates its cells."""
    rng, shape, minb = _bounds(data_type, voxel_size)
    lib = _lib.get_lib()
    output = tuple()
    for data in args:
        is_np = len(data) > 0 and isinstance(data[0], np.ndarray)
        clouds = [_dev(p)[0] for p in data]
        vol = torch.zeros(shape, dtype=torch.float32, device=_lib.compute_device())
        if clouds:
            lens = [int(c.shape[0]) for c in clouds]
            stride = min(int(c.shape[1]) for c in clouds)
            cat = torch.cat([c[:, :stride] for c in clouds], dim=0).contiguous()
            off = torch.tensor(np.concatenate([[0], np.cumsum(lens)]), dtype=torch.int32, device=cat.device)
            words = (shape[0] * shape[1] + 31) // 32
            chunk = max(1, min(len(clouds), 65535, (1 << 31) // (4 * words)))     # bitmap scratch <= 2 GiB per launch
            for c0 in range(0, len(clouds), chunk):
                c1 = min(c0 + chunk, len(clouds))
                bm = torch.empty((c1 - c0) * words, dtype=torch.int32, device=cat.device)
                lib.bev_occupancy_sum(cat.data_ptr(), off[c0:].data_ptr(), c1 - c0, max(lens[c0:c1]), stride,
                                      float(rng[0][0]), float(rng[0][1]), float(rng[1][0]), float(rng[1][1]),
                                      float(voxel_size), minb[0], minb[1], shape[0], shape[1], bm.data_ptr(),
                                      vol.data_ptr(), _stream())
        output += (_ret(vol, is_np),)
    return output


def pcd2voxel_full(data_type, *args):
    """metric_utils.py:170-199: per cloud a dense [X,Y,Z] fp32 occupancy volume (1 where a point falls)."""
    rng, shape, minb = _bounds(data_type, VOXEL_SIZE, dims=3)
    lib = _lib.get_lib()
    lohi = (C.c_float * 6)(*[float(v) for r in rng for v in r])
    mb = (C.c_int32 * 3)(*minb)
    dm = (C.c_int32 * 3)(*shape)
    output = tuple()
    for data in args:
        volume_list = []
        for pcd in data:
            pts, is_np = _dev(pcd)
            vol = torch.empty(shape, dtype=torch.float32, device=pts.device)
            lib.voxel_occupancy(pts.data_ptr(), pts.shape[0], pts.stride(0), lohi, float(VOXEL_SIZE), mb, dm, vol.data_ptr(),
                                _stream())
            volume_list.append(_ret(vol, is_np))
        output += (volume_list,)
    return output


def pcd2bev_bin(data_type, *args, voxel_size=0.5):
    """metric_utils.py:259-283: unique BEV cells of every cloud (ravel-hash order), normalised to [0,1)."""
    rng, shape, minb = _bounds(data_type, voxel_size)
    output = tuple()
    for data in args:
        pcd_list = []
        for pcd in data:
            pts, is_np = _dev(pcd)
            xy = _bev_filter(pts, rng)
            if xy.shape[0] == 0:
                pcd_list.append(_ret(torch.empty(0, 2, device=pts.device), is_np))
                continue
            voxel, mm = _quantize(xy, (voxel_size, voxel_size), div_f32=True)
            uniq, _, _ = _unique(voxel, mm, want_inverse=False)
            cells = (uniq.double() - torch.tensor(minb, dtype=torch.float64, device=uniq.device)) / \
                torch.tensor(shape, dtype=torch.float64, device=uniq.device)
            pcd_list.append(_ret(cells.float(), is_np))
        output += (pcd_list,)
    return output


def bev_sample(data_type, *args, voxel_size=0.5):
    """metric_utils.py:286-306: the first point (x, y) of every occupied BEV cell, in ravel-hash order."""
    rng, _, _ = _bounds(data_type, voxel_size)
    output = tuple()
    for data in args:
        pcd_list = []
        for pcd in data:
            pts, is_np = _dev(pcd)
            xy = _bev_filter(pts, rng)
            if xy.shape[0] == 0:
                pcd_list.append(_ret(xy, is_np))
                continue
            voxel, mm = _quantize(xy, (voxel_size, voxel_size), div_f32=True)
            _, idx, _ = _unique(voxel, mm, want_inverse=False)
            pcd_list.append(_ret(xy[idx], is_np))
        output += (pcd_list,)
    return output
