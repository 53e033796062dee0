"""ctypes wrapper of the metrics part of oracle/liboracle.so (oracle/metrics_ops.c).  TEST ORACLE ONLY."""
import ctypes as C
import math

import numpy as np

from .lidar_ops import _p, lib

VOXEL_SIZE = 0.05
DATA_CONFIG = {'64': {'x': [-50, 50], 'y': [-50, 50], 'z': [-3, 1]},
               '32': {'x': [-30, 30], 'y': [-30, 30], 'z': [-3, 6]}}


def pcd2range(pcd, size, fov, depth_range, feature=None, fill=-1.0):
    pts = np.ascontiguousarray(pcd[:, :3], dtype=np.float32)
    H, W = size
    rng = np.zeros((H, W), np.float32)
    win = np.zeros((H, W), np.int32)
    pf = np.zeros((H, W), np.float32) if feature is not None else None
    ft = np.ascontiguousarray(feature, dtype=np.float32) if feature is not None else None
    lib().oracle_pcd2range(_p(pts), _p(ft) if ft is not None else None, pts.shape[0], H, W, C.c_float(fov[0]),
                           C.c_float(fov[1]), C.c_float(depth_range[0]), C.c_float(depth_range[1]), C.c_float(fill),
                           _p(rng), _p(pf) if pf is not None else None, _p(win))
    return rng, pf, win


def range2xyz(img, fov, depth_range, depth_scale=0.0, log_scale=False):
    im = np.ascontiguousarray(img, dtype=np.float32)
    H, W = im.shape
    xyz = np.zeros((3, H, W), np.float64)
    lib().oracle_range2xyz(_p(im), H, W, C.c_float(fov[0]), C.c_float(fov[1]), C.c_float(depth_range[0]),
                           C.c_float(depth_range[1]), C.c_float(depth_scale), int(log_scale), _p(xyz))
    return xyz


def quantize(coords, voxel_size, div_f32):
    c = np.ascontiguousarray(coords)
    assert c.dtype in (np.float32, np.float64)
    M, D = c.shape
    vs = np.asarray(list(voxel_size), np.float64)
    out = np.zeros((M, D), np.int32)
    lib().oracle_quantize(_p(c), int(c.dtype == np.float64), M, D, D, _p(vs), int(div_f32), _p(out))
    return out


def sparse_quantize(voxel):
    """int32 [M,D] -> (keys uint64 [M], uniq [U,D], indices int64 [U], inverse int64 [M])"""
    v = np.ascontiguousarray(voxel, dtype=np.int32)
    M, D = v.shape
    keys = np.zeros(M, np.uint64)
    uniq = np.zeros((M, D), np.int32)
    idx = np.zeros(M, np.int64)
    inv = np.zeros(M, np.int64)
    n = lib().oracle_sparse_quantize(_p(v), M, D, _p(keys), _p(uniq), _p(idx), _p(inv))
    return keys, uniq[:n], idx[:n], inv


def bounds(data_type, voxel_size, dims=2):
    cfg = DATA_CONFIG[data_type]
    rng = [cfg['x'], cfg['y'], cfg['z']][:dims]
    shape = tuple(math.ceil((r[1] - r[0]) / voxel_size) for r in rng)
    minb = tuple(math.ceil(r[0] / voxel_size) for r in rng)
    return rng, shape, minb


def bev_sum(data_type, clouds, voxel_size=VOXEL_SIZE):
    rng, shape, minb = bounds(data_type, voxel_size)
    stride = min(c.shape[1] for c in clouds)
    cat = np.ascontiguousarray(np.concatenate([c[:, :stride] for c in clouds]), dtype=np.float32)
    off = np.concatenate([[0], np.cumsum([len(c) for c in clouds])]).astype(np.int32)
    lo = np.array([r[0] for r in rng], np.float32)
    hi = np.array([r[1] for r in rng], np.float32)
    vol = np.zeros(shape, np.float32)
    lib().oracle_bev_sum(_p(cat), _p(off), len(clouds), stride, _p(lo), _p(hi), C.c_float(voxel_size),
                         _p(np.array(minb, np.int32)), _p(np.array(shape, np.int32)), _p(vol))
    return vol


def voxel_full(data_type, pcd):
    rng, shape, minb = bounds(data_type, VOXEL_SIZE, dims=3)
    pts = np.ascontiguousarray(pcd, dtype=np.float32)
    lo = np.array([r[0] for r in rng], np.float32)
    hi = np.array([r[1] for r in rng], np.float32)
    vol = np.zeros(shape, np.float32)
    lib().oracle_voxel_full(_p(pts), pts.shape[0], pts.shape[1], _p(lo), _p(hi), C.c_float(VOXEL_SIZE),
                            _p(np.array(minb, np.int32)), _p(np.array(shape, np.int32)), _p(vol))
    return vol
