"""Pin oracle/lidar_ops.c to the reference's own NumPy projection (tests/golden/projection.npz, produced by
tests/golden/make_golden_lidar.py from lidargen/dataset/transforms_3d/common.py:26-91), and check the
points-in-boxes / voxel-index restatements against independent NumPy arithmetic."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import pytest  # noqa: E402
from make_golden_lidar import synth_sweep  # noqa: E402
from make_golden_roiaware import synth_box_points, synth_boxes  # noqa: E402
from oracle import lidar_ops as LO  # noqa: E402

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "projection.npz"))


def test_projection_bins_and_zbuffer_match_reference():
    for seed in (0, 1, 2):
        pts = synth_sweep(seed)
        img, grid, win = LO.range_project(pts)
        assert np.array_equal(grid, GOLD[f"grid_{seed}"].astype(np.int32))
        assert np.array_equal(win, GOLD[f"win_{seed}"])
        assert abs(img.astype(np.float64).sum() - GOLD[f"imgsum_{seed}"][0]) < 1e-6 * abs(GOLD[f"imgsum_{seed}"][0])
        assert int(img[..., 5].sum()) == int(GOLD[f"imgsum_{seed}"][1])


def test_projection_edge_cases():
    # a point at the origin, points on the -x axis (azimuth wrap), duplicates (depth ties)
    pts = np.array([[0, 0, 0, 1], [0, 5, 0, 2], [-5, 0, 0, 3], [-5, -0.0, 0, 4], [0, 5, 0, 9]], np.float32)
    img, grid, win = LO.range_project(pts)
    assert grid[0].tolist() == [8, 512]           # origin: elevation 0 -> row floor((1 - 30/40) * 32) = 8, azimuth 0
    assert grid[1].tolist() == [8, 256]
    assert win[8, 256] == 4                       # equal depth: highest index wins
    assert grid[2, 1] in (0, 1023) and grid[3, 1] in (0, 1023)
    assert img[8, 512, 5] == 0.0                  # depth 0 < min_depth -> mask 0 (but it still owns the pixel)


ROI_GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "roiaware.npz"))


def test_points_in_boxes_oracle_matches_reference_cpp_golden():
    """tests/golden/roiaware.npz = outputs of the reference's own roiaware_pool3d.cpp (oracle/_ref); a third of the
    points lie within 2e-6 m of a face, so this pins the rounding of the restatement, not just its logic."""
    for seed in (0, 1):
        boxes = synth_boxes(seed)
        pts = synth_box_points(seed, boxes)
        got = LO.points_in_boxes(pts, boxes)
        want = np.unpackbits(ROI_GOLD[f"bits_{seed}"], axis=1)[:, :pts.shape[0]].astype(np.int32)
        assert int(want.sum()) == int(ROI_GOLD[f"count_{seed}"][0]) > 5000
        assert np.array_equal(got, want)
        big = boxes.copy(); big[:, 3:6] += 0.2
        want_big = np.unpackbits(ROI_GOLD[f"bits_big_{seed}"], axis=1)[:, :pts.shape[0]].astype(np.int32)
        assert np.array_equal(LO.points_in_boxes(pts, big), want_big)


@pytest.mark.skipif(not LO.ref_cpu_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_points_in_boxes_oracle_matches_reference_cpp_live():
    """the compiled reference itself on fresh seeds (build container, or a box that received the prebuilt oracle/_ref)"""
    for seed in (7, 8, 9):
        boxes = synth_boxes(seed)
        pts = synth_box_points(seed, boxes, per_box=4000)
        assert np.array_equal(LO.points_in_boxes(pts, boxes), LO.ref_points_in_boxes_cpu(pts, boxes))
    # empty / degenerate inputs the reference accepts
    boxes = synth_boxes(3, 1)
    assert LO.ref_points_in_boxes_cpu(np.zeros((0, 3), np.float32), boxes).shape == (1, 0)
    flat = boxes.copy(); flat[:, 5] = 0.0                      # zero-height box: only z == cz passes the first test
    p = np.array([[flat[0, 0], flat[0, 1], flat[0, 2]], [flat[0, 0], flat[0, 1], flat[0, 2] + 1e-3]], np.float32)
    assert np.array_equal(LO.points_in_boxes(p, flat), LO.ref_points_in_boxes_cpu(p, flat))


def _boxes(rs, n):
    c = rs.uniform(-30, 30, (n, 3)); c[:, 2] = rs.uniform(-2, 0, n)
    s = rs.uniform(0.5, 6.0, (n, 3))
    yaw = rs.uniform(-np.pi, np.pi, (n, 1))
    return np.concatenate([c, s, yaw], 1).astype(np.float32)


def test_points_in_boxes_against_numpy_float64():
    rs = np.random.RandomState(0)
    boxes = _boxes(rs, 12)
    # points concentrated around the boxes so that a fair share is inside
    pts = (boxes[rs.randint(0, 12, 5000), :3] + rs.normal(0, 2.0, (5000, 3))).astype(np.float32)
    got = LO.points_in_boxes(pts, boxes)
    d = pts[None, :, :].astype(np.float64) - boxes[:, None, :3]
    ca, sa = np.cos(-boxes[:, 6].astype(np.float64))[:, None], np.sin(-boxes[:, 6].astype(np.float64))[:, None]
    lx = d[..., 0] * ca - d[..., 1] * sa
    ly = d[..., 0] * sa + d[..., 1] * ca
    m = 1e-2
    inside = (np.abs(d[..., 2]) <= boxes[:, None, 5] / 2.0) & (np.abs(lx) < boxes[:, None, 3] / 2.0 + m) & \
             (np.abs(ly) < boxes[:, None, 4] / 2.0 + m)
    margin = np.minimum(np.abs(np.abs(lx) - boxes[:, None, 3] / 2.0 - m), np.abs(np.abs(ly) - boxes[:, None, 4] / 2.0 - m))
    margin = np.minimum(margin, np.abs(np.abs(d[..., 2]) - boxes[:, None, 5] / 2.0))
    clear = margin > 1e-4                      # away from the faces both must agree
    assert got.sum() > 50
    assert np.array_equal(got[clear], inside[clear].astype(np.int32))


def test_voxel_index_encoding():
    box = np.array([[0, 0, 0, 4, 2, 2, 0.0]], np.float32)
    pts = np.array([[-1.99, -0.99, -0.99], [1.99, 0.99, 0.99], [0.0, 0.0, 0.0], [5, 0, 0]], np.float32)
    code = LO.voxel_index(pts, box, (14, 14, 14))[0]
    assert code[0] == 0
    assert code[1] == (13 << 16) + (13 << 8) + 13
    # fp32 arithmetic of the reference kernel: int((0 + d/2) / (d/14)) with d/14 rounded up -> 6, not 7
    f = np.float32
    exp = [int((f(0) + f(d) / f(2)) / (f(d) / f(14))) for d in (4, 2, 2)]
    assert exp == [6, 6, 6]
    assert code[2] == (exp[0] << 16) + (exp[1] << 8) + exp[2]
    assert code[3] == -1
    first = LO.points_in_boxes_first(pts[None], np.concatenate([box, box])[None])
    assert first[0].tolist() == [0, 0, 0, -1]
