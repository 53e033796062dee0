"""GPU parity of the point-cloud ops (bit-exact integer outputs) through the reference-named Python API,
which calls the C-ABI: against the reference's golden vectors and against oracle/lidar_ops.c."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
from make_golden_lidar import synth_sweep  # noqa: E402
from oracle import lidar_ops as LO  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "projection.npz"))


def test_projection_vs_reference_golden_and_oracle():
    from lidarcrafter_b200 import ops
    for seed in (0, 1, 2):
        pts = synth_sweep(seed)
        img, grid = ops.load_points_as_images(points=pts, scan_unfolding=False, H=32, W=1024, return_grid=True)
        assert np.array_equal(grid, GOLD[f"grid_{seed}"].astype(np.int32)), "range bins differ from the reference"
        o_img, o_grid, o_win = LO.range_project(pts)
        assert np.array_equal(grid, o_grid)
        assert np.array_equal(img, o_img), "range image differs from the oracle (bit-exact expected)"
        win = GOLD[f"win_{seed}"]
        occ = win >= 0
        assert np.array_equal(img[occ][:, :4], pts[win[occ]])


def test_projection_batched_ragged_and_ties():
    from lidarcrafter_b200 import ops
    frames = [synth_sweep(5)[:20000], synth_sweep(6)[:20000]]
    frames[1][1000:2000] = frames[1][0:1000]          # exact duplicates -> depth ties
    batch = np.stack(frames)
    img = ops.load_points_as_images(points=batch, H=32, W=1024)
    for f in range(2):
        o_img, _, _ = LO.range_project(frames[f])
        assert np.array_equal(img[f], o_img)
    one = np.array([[3, 4, 0.5, 7]], np.float32)
    img1 = ops.load_points_as_images(points=one, H=32, W=1024)
    assert (img1[..., 4] > 0).sum() == 1 and np.array_equal(img1, LO.range_project(one)[0])


def _boxes(rs, n):
    c = rs.uniform(-30, 30, (n, 3)); c[:, 2] = rs.uniform(-2, 0, n)
    s = rs.uniform(0.5, 6.0, (n, 3))
    yaw = rs.uniform(-np.pi, np.pi, (n, 1))
    return np.concatenate([c, s, yaw], 1).astype(np.float32)


def test_points_in_boxes_and_voxel_index_bit_exact():
    from lidarcrafter_b200 import ops
    rs = np.random.RandomState(0)
    boxes = _boxes(rs, 13)
    pts = synth_sweep(3)[:, :3].copy()
    pts[:6000] = (boxes[rs.randint(0, 13, 6000), :3] + rs.normal(0, 1.5, (6000, 3))).astype(np.float32)
    got = ops.points_in_boxes_cpu(pts, boxes.copy())
    big = boxes.copy(); big[:, 3:6] += np.float32(0.2)
    assert np.array_equal(got, LO.points_in_boxes(pts, big))
    assert got.sum() > 100
    first = ops.points_in_boxes_gpu(torch.from_numpy(pts[None]), torch.from_numpy(boxes[None])).cpu().numpy()
    assert np.array_equal(first, LO.points_in_boxes_first(pts[None], boxes[None]))
    code = ops.voxel_index(pts, boxes, (14, 14, 14))
    assert np.array_equal(code, LO.voxel_index(pts, boxes, (14, 14, 14)))
    assert (code >= 0).sum() > 100


def test_depth_to_xyz_vs_numpy():
    import lidarcrafter_b200 as L
    ang = L.get_linear_ray_angles(32, 1024, 10, -30)
    lu = L.LiDARUtility((32, 1024), "log_depth", 1.45, 80.0, ang).cuda()
    x = torch.rand(2, 1, 32, 1024, generator=torch.Generator().manual_seed(0)) * 2.2 - 1.1
    depth, xyz = lu.to_xyz_from_normalized(x.cuda())
    m_ref, xyz_ref = LO.depth_to_xyz(x.numpy(), ang.numpy())
    assert np.allclose(depth.cpu().numpy(), m_ref, rtol=2e-5, atol=1e-4)
    assert np.allclose(xyz.cpu().numpy(), xyz_ref, rtol=2e-5, atol=2e-4)
    # the fused kernel equals the reference's three-call chain
    chain = lu.to_xyz(lu.revert_depth(lu.denormalize(x.cuda())))
    assert torch.allclose(chain, xyz, rtol=2e-5, atol=2e-4)


def test_rollout_glue_on_device():
    """delete_fg_points / extract_object_points / get_next_frame_points (pipe_related.py:54-68,243-288) against a
    NumPy restatement built on the C oracle."""
    from lidarcrafter_b200 import rollout
    rs = np.random.RandomState(1)
    boxes = _boxes(rs, 6)
    pts = synth_sweep(4)
    pts[:3000, :3] = (boxes[rs.randint(0, 6, 3000), :3] + rs.normal(0, 1.0, (3000, 3))).astype(np.float32)
    tp, tb = torch.from_numpy(pts).cuda(), torch.from_numpy(boxes).cuda()
    big = boxes.copy(); big[:, 3:6] += np.float32(0.2)
    m = LO.points_in_boxes(pts[:, :3], big)
    bg = rollout.delete_fg_points(tp, tb).cpu().numpy()
    assert np.array_equal(bg, pts[m.sum(0) == 0])
    objs, inten = rollout.extract_object_points(tp, tb)
    for k in range(6):
        sel = pts[m[k] > 0]
        assert objs[k].shape[0] == sel.shape[0] and np.array_equal(inten[k].cpu().numpy(), sel[:, 3])
        c, s = np.cos(-boxes[k, 6]), np.sin(-boxes[k, 6])
        ref = (sel[:, :3] - boxes[k, :3]) @ np.array([[c, s, 0], [-s, c, 0], [0, 0, 1]], np.float32)
        assert np.allclose(objs[k].cpu().numpy(), ref, atol=1e-4)
    T = rollout.compute_inter_frame_transforms(np.array([[0.02, 0.5]]))[0]
    nxt = rollout.get_next_frame_points(torch.from_numpy(bg).cuda(), objs, inten, tb, T).cpu().numpy()
    # oracle: warp in fp64, project with the C oracle, drop masked / empty pixels, paste the objects
    h = np.concatenate([bg[:, :3].astype(np.float64), np.ones((len(bg), 1))], 1)
    w = (T @ h.T).T
    w[:, 3] = bg[:, 3]
    img, _, _ = LO.range_project(w.astype(np.float32))
    img = img * img[..., 5:6]
    p = img[..., :4].reshape(-1, 4)
    p = p[np.linalg.norm(p[:, :3], axis=1) > 1e-2]
    assert nxt.shape[0] == p.shape[0] + sum(o.shape[0] for o in objs)
    assert np.array_equal(nxt[:p.shape[0]], p)
