"""GPU parity of the evaluation-side point-cloud ops (lidarcrafter_b200.metric_utils -> C-ABI -> metrics_ops.cu):
bit-exact against the reference's golden vectors (tests/golden/metrics.npz) and against oracle/metrics_ops.c, plus
size-independent properties at sizes the sort-based oracle would not finish quickly."""
import os
import sys

import numpy as np
import pytest
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from make_golden_lidar import synth_sweep  # noqa: E402
from make_golden_metrics import KITTI, NUSC, digest, small_cloud  # noqa: E402
from oracle import metrics_ops as MO  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(HERE, "golden", "metrics.npz"))


def same(a, key):
    return np.array_equal(digest(a), GOLD[key])


def test_pcd2range_vs_reference_golden_and_oracle():
    from lidarcrafter_b200 import metric_utils as MU
    for seed in (0, 1):
        sw = synth_sweep(seed)
        pts, inten = sw[:, :3].copy(), sw[:, 3].copy()
        r, f = MU.pcd2range(pts, remission=inten, **NUSC)
        assert r.dtype == np.float32 and np.array_equal(r, GOLD[f"range_{seed}"]), "range image differs from the reference"
        assert same(f, f"feat_digest_{seed}")
        assert same(MU.pcd2range(pts, **KITTI)[0], f"range_kitti_digest_{seed}")
        lab = (np.arange(len(pts)) % 17).astype(np.int64)
        assert same(MU.pcd2range(pts, labels=lab, **NUSC)[1], f"label_digest_{seed}")
    # batched (device tensors in, device tensors out), ragged content, depth ties, an empty frame
    a, b = synth_sweep(3)[:20000, :3].copy(), synth_sweep(4)[:20000, :3].copy()
    b[1000:2000] = b[0:1000]
    b[5000:] = 0.0                                           # depth 0 -> rejected by the strict range test
    batch = torch.from_numpy(np.stack([a, b, np.zeros_like(a)])).cuda()
    feat = torch.arange(20000, dtype=torch.float32).repeat(3, 1).cuda()
    r, f = MU.pcd2range(batch, remission=feat, **NUSC)
    for k, c in enumerate((a, b, np.zeros_like(a))):
        ro, fo, _ = MO.pcd2range(c, NUSC["size"], NUSC["fov"], NUSC["depth_range"], feature=np.arange(20000, dtype=np.float32))
        assert np.array_equal(r[k].cpu().numpy(), ro) and np.array_equal(f[k].cpu().numpy(), fo)
    assert (r[2] == -1).all()


def test_range2xyz_and_preprocess_range():
    from lidarcrafter_b200 import metric_utils as MU
    for seed in (0, 1):
        xyz = MU.range2xyz(GOLD[f"range_{seed}"], log_scale=False, **NUSC)
        assert xyz.dtype == np.float64 and xyz.shape == (3, 32, 1024)
        assert np.allclose(xyz[:, ::2, ::16], GOLD[f"xyz_sub_{seed}"], rtol=1e-12, atol=1e-11)
        assert abs(np.abs(xyz).sum() - GOLD[f"xyz_abs_sum_{seed}"][0]) < 1e-9 * GOLD[f"xyz_abs_sum_{seed}"][0]
        pr = MU.preprocess_range(synth_sweep(seed)[:, :3].copy(), **KITTI)
        assert pr.shape == (4, 64, 1024) and same(pr[0].astype(np.float32), f"prep_range_digest_{seed}")
        assert abs(np.abs(pr).sum() - GOLD[f"prep_range_abs_sum_{seed}"][0]) < 1e-9 * GOLD[f"prep_range_abs_sum_{seed}"][0]
    rk = MU.pcd2range(synth_sweep(0)[:, :3].copy(), **KITTI)[0]
    logimg = (np.log2(np.maximum(rk, 0) + 1) / 6).astype(np.float32)
    xyzk = MU.range2xyz(logimg, **KITTI)
    assert np.allclose(xyzk[:, ::4, ::16], GOLD["xyz_log_sub"], rtol=1e-6, atol=1e-6)   # exp2f: tolerance 1e-6 relative


def test_sparse_quantize_vs_reference_golden():
    from lidarcrafter_b200 import metric_utils as MU
    for seed in (0, 1):
        pts = synth_sweep(seed)[:, :3].copy()
        c, i, inv = MU.sparse_quantize(pts, 0.05, return_index=True, return_inverse=True)
        assert c.dtype == np.int32 and i.dtype == np.int64 and inv.dtype == np.int64
        assert len(c) == int(GOLD[f"sq_n_{seed}"][0])
        assert same(c, f"sq_coords_digest_{seed}") and same(i, f"sq_index_digest_{seed}")
        assert same(inv, f"sq_inverse_digest_{seed}")
        vox = np.floor(pts / np.array([0.05] * 3)).astype(np.int32)
        assert same(MU.ravel_hash(vox), f"hash_digest_{seed}")
    sc = small_cloud(5)
    c, i, inv = MU.sparse_quantize(sc, (0.2, 0.2, 0.1), return_index=True, return_inverse=True)
    assert np.array_equal(c, GOLD["small_coords"]) and np.array_equal(i, GOLD["small_index"])
    assert np.array_equal(inv, GOLD["small_inverse"])
    c2, i2 = MU.sparse_quantize(sc[:, :2].astype(np.float64), 0.25, return_index=True)
    assert np.array_equal(c2, GOLD["small2d_coords"]) and np.array_equal(i2, GOLD["small2d_index"])
    only = MU.sparse_quantize(sc, 0.5)
    assert isinstance(only, np.ndarray) and only.shape[1] == 3
    # a single point / all points in one voxel
    one = MU.sparse_quantize(np.array([[1.0, 2.0, 3.0]], np.float32), 1, return_index=True, return_inverse=True)
    assert one[0].tolist() == [[1, 2, 3]] and one[1].tolist() == [0] and one[2].tolist() == [0]
    rep = MU.sparse_quantize(np.tile(np.array([[0.4, 0.4, 0.4]], np.float32), (1000, 1)), 1, return_index=True, return_inverse=True)
    assert rep[0].tolist() == [[0, 0, 0]] and rep[1].tolist() == [0] and (rep[2] == 0).all()


def test_sparse_quantize_properties_at_scale():
    """2 M points over a 100 m x 100 m x 8 m grid at 5 cm (bounding grid 6.4e8 cells = 80 MB bitmap): inverse rebuilds
    the input, keys strictly increase, indices are first occurrences, and it agrees with torch.unique on the keys."""
    from lidarcrafter_b200 import metric_utils as MU
    g = torch.Generator(device="cuda").manual_seed(0)
    pts = (torch.rand(2_000_000, 3, device="cuda", generator=g) - 0.5) * torch.tensor([100.0, 100.0, 8.0], device="cuda")
    pts[500_000:1_000_000] = pts[:500_000]                      # guaranteed duplicates
    c, i, inv = MU.sparse_quantize(pts, 0.05, return_index=True, return_inverse=True)
    vox = torch.floor(pts.double() / 0.05).to(torch.int32)
    assert torch.equal(c[inv], vox)
    assert torch.equal(vox[i], c)
    keys = MU.ravel_hash(vox)
    assert bool((keys[i][1:] > keys[i][:-1]).all())
    ku, kinv = torch.unique(keys, return_inverse=True)
    assert ku.numel() == c.shape[0] and torch.equal(kinv, inv)
    first = torch.full((c.shape[0],), 2**62, dtype=torch.int64, device="cuda")
    first.scatter_reduce_(0, inv, torch.arange(pts.shape[0], device="cuda"), reduce="amin")
    assert torch.equal(first, i)


def test_bev_and_voxel_grids_vs_reference_golden():
    from lidarcrafter_b200 import metric_utils as MU
    for seed in (0, 1):
        pts = synth_sweep(seed)[:, :3].copy()
        clouds = [pts, pts[:5000] + np.float32(0.3), small_cloud(seed)]
        bs = MU.pcd2bev_sum('32', clouds)[0]
        nz = np.flatnonzero(bs)
        assert bs.shape == (1200, 1200) and same(np.stack([nz, bs.ravel()[nz].astype(np.int64)]), f"bev_sum_digest_{seed}")
        assert [bs.sum(), bs.max()] == GOLD[f"bev_sum_total_{seed}"].tolist()
        assert np.array_equal(bs, MO.bev_sum('32', clouds))
        vf = MU.pcd2voxel_full('32', [pts])[0][0]
        assert vf.shape == (1200, 1200, 180) and vf.sum() == GOLD[f"voxel_full_total_{seed}"][0]
        assert same(np.flatnonzero(vf), f"voxel_full_digest_{seed}")
        bb = MU.pcd2bev_bin('32', [pts, small_cloud(seed)])[0]
        assert [len(b) for b in bb] == GOLD[f"bev_bin_n_{seed}"].tolist()
        assert bb[0].dtype == np.float32 and same(np.concatenate(bb), f"bev_bin_digest_{seed}")
        sm = MU.bev_sample('32', [pts, small_cloud(seed)])[0]
        assert same(np.concatenate(sm), f"bev_sample_digest_{seed}")
    # two data sets in one call (reference, samples) and a cloud entirely outside the range
    far = np.full((10, 3), 100.0, np.float32)
    a, b = MU.pcd2bev_sum('32', [far], [synth_sweep(2)[:, :3].copy(), far])
    assert a.sum() == 0 and b.sum() > 1000
    assert MU.pcd2bev_bin('32', [far])[0][0].shape == (0, 2)
