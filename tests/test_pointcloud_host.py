"""Host mirrors of the point-cloud interfaces (lidarcrafter_b200/ops.py, metric_utils.py) without a GPU: the same
Python that drives the CUDA kernels runs here through the C-ABI emulator (whose point-cloud entries are the C oracle),
and must reproduce the reference's golden vectors -- batching, bounds, list / dtype handling, workspace protocol."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from abi_emulator import EmulatedLib  # noqa: E402
from make_golden_lidar import synth_sweep  # noqa: E402
from make_golden_metrics import KITTI, NUSC, digest, small_cloud  # noqa: E402
from make_golden_roiaware import synth_box_points, synth_boxes  # noqa: E402
from lidarcrafter_b200 import _lib  # noqa: E402

GOLD = np.load(os.path.join(HERE, "golden", "metrics.npz"))
PROJ = np.load(os.path.join(HERE, "golden", "projection.npz"))
ROI = np.load(os.path.join(HERE, "golden", "roiaware.npz"))


@pytest.fixture()
def emu():
    lib = EmulatedLib()
    _lib.set_test_lib(lib)
    yield lib
    _lib.set_test_lib(None)


def same(a, key):
    return np.array_equal(digest(a), GOLD[key])


def test_ops_host_logic(emu):
    from lidarcrafter_b200 import ops
    pts = synth_sweep(0)
    img, grid = ops.load_points_as_images(points=pts, H=32, W=1024, return_grid=True)
    assert img.shape == (32, 1024, 6) and np.array_equal(grid, PROJ["grid_0"].astype(np.int32))
    batch = ops.load_points_as_images(points=np.stack([pts[:20000], synth_sweep(1)[:20000]]), H=32, W=1024)
    assert batch.shape == (2, 32, 1024, 6)
    with pytest.raises(NotImplementedError):
        ops.load_points_as_images(points=pts, scan_unfolding=True)
    boxes = synth_boxes(0)
    p = synth_box_points(0, boxes)
    got = ops.points_in_boxes_cpu(p, boxes.copy())          # numpy in -> numpy out, boxes enlarged by 0.2 m inside
    assert isinstance(got, np.ndarray)
    assert np.array_equal(got, np.unpackbits(ROI["bits_big_0"], axis=1)[:, :p.shape[0]].astype(np.int32))
    tb = torch.from_numpy(boxes.copy())
    ops.points_in_boxes_cpu(torch.from_numpy(p), tb)
    assert torch.allclose(tb[:, 3:6], torch.from_numpy(boxes[:, 3:6]) + 0.2)   # the reference mutates the caller's tensor
    first = ops.points_in_boxes_gpu(torch.from_numpy(p[None]), torch.from_numpy(boxes[None]))
    assert first.shape == (1, p.shape[0]) and int(first.max()) == 12
    assert ops.voxel_index(p, boxes, 14).shape == (13, p.shape[0])
    assert emu.calls.count("points_in_boxes") == 2


def test_metric_utils_host_logic(emu):
    from lidarcrafter_b200 import metric_utils as MU
    pts = synth_sweep(0)[:, :3].copy()
    r, f = MU.pcd2range(pts, remission=synth_sweep(0)[:, 3].copy(), **NUSC)
    assert np.array_equal(r, GOLD["range_0"]) and same(f, "feat_digest_0")
    assert MU.pcd2range(pts, **NUSC)[1] is None
    xyz = MU.range2xyz(r, log_scale=False, **NUSC)
    assert xyz.dtype == np.float64 and np.allclose(xyz[:, ::2, ::16], GOLD["xyz_sub_0"], rtol=1e-12, atol=1e-12)
    pr = MU.preprocess_range(pts, **KITTI)
    assert pr.shape == (4, 64, 1024) and same(pr[0].astype(np.float32), "prep_range_digest_0")
    c, i, inv = MU.sparse_quantize(pts, 0.05, return_index=True, return_inverse=True)
    assert same(c, "sq_coords_digest_0") and same(i, "sq_index_digest_0") and same(inv, "sq_inverse_digest_0")
    sc = small_cloud(5)
    c2, i2 = MU.sparse_quantize(sc[:, :2].astype(np.float64), 0.25, return_index=True)      # fp64 coordinates, 2-D
    assert np.array_equal(c2, GOLD["small2d_coords"]) and np.array_equal(i2, GOLD["small2d_index"])
    assert isinstance(MU.sparse_quantize(sc, (0.2, 0.2, 0.1)), np.ndarray)                  # single output -> bare array
    vox = np.floor(pts / np.array([0.05] * 3)).astype(np.int32)
    assert same(MU.ravel_hash(vox), "hash_digest_0")
    clouds = [pts, pts[:5000] + np.float32(0.3), small_cloud(0)]
    far = np.full((10, 3), 100.0, np.float32)
    ref_sum, smp_sum = MU.pcd2bev_sum('32', clouds, [far])                                  # two data sets, one call
    nz = np.flatnonzero(ref_sum)
    assert same(np.stack([nz, ref_sum.ravel()[nz].astype(np.int64)]), "bev_sum_digest_0") and smp_sum.sum() == 0
    bb = MU.pcd2bev_bin('32', [pts, small_cloud(0), far])[0]
    assert [len(b) for b in bb[:2]] == GOLD["bev_bin_n_0"].tolist() and bb[2].shape == (0, 2)
    assert same(np.concatenate(bb[:2]), "bev_bin_digest_0")
    assert same(np.concatenate(MU.bev_sample('32', [pts, small_cloud(0)])[0]), "bev_sample_digest_0")
    kept = MU.preprocess_pcd(pts, depth_range=[1.0, 45.0])
    d = np.linalg.norm(pts, axis=1)
    assert kept.shape[0] == int(((d > 1.0) & (d < 45.0)).sum())


def test_sparse_quantize_refuses_absurd_extents(emu):
    from lidarcrafter_b200 import metric_utils as MU
    pts = np.array([[0, 0, 0], [1e6, 1e6, 1e5]], np.float32)       # bounding grid of 2e7 x 2e7 x 2e6 cells at 5 cm
    with pytest.raises(_lib.B200LidarError):
        MU.sparse_quantize(pts, 0.05)
