"""Golden vectors for points_in_boxes_cpu from the reference's OWN C++ (build container only).

    python tests/golden/make_golden_roiaware.py

oracle/_ref/libref_roiaware_cpu.so is lidargen/ops/roiaware_pool3d/src/roiaware_pool3d.cpp compiled unmodified from
/root/reference (oracle/Makefile).  Inputs are seeded (synth_box_points); the (N, M) 0/1 result is stored bit-packed.
A third of the points sit within ~2e-6 m of a box face (incl. the 1e-2 MARGIN of the x/y faces), which is where a
restatement with different rounding would disagree.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))


def synth_boxes(seed: int, n: int = 13):
    rs = np.random.RandomState(seed)
    b = np.zeros((n, 7), np.float32)
    b[:, :2] = rs.uniform(-40, 40, (n, 2))
    b[:, 2] = rs.uniform(-2, 0, n)
    b[:, 3:6] = rs.uniform(0.5, 6.0, (n, 3))
    b[:, 6] = rs.uniform(-np.pi, np.pi, n)
    return b


def synth_box_points(seed: int, boxes: np.ndarray, per_box: int = 1500, margin: float = 1e-2):
    """points in / around every box; the first third of each group is snapped onto a face (+- 2e-6 m jitter)"""
    rs = np.random.RandomState(seed + 1000)
    out = []
    for bx in boxes.astype(np.float64):
        c, d, rz = bx[:3], bx[3:6], bx[6]
        loc = rs.uniform(-0.6, 0.6, (per_box, 3)) * d
        k = per_box // 3
        ax = rs.randint(0, 3, k)
        sgn = rs.choice([-1.0, 1.0], k)
        mg = np.where(ax == 2, 0.0, margin)
        loc[np.arange(k), ax] = sgn * (d[ax] / 2 + mg) + rs.normal(0, 2e-6, k)
        ca, sa = np.cos(rz), np.sin(rz)
        out.append(np.stack([loc[:, 0] * ca - loc[:, 1] * sa + c[0], loc[:, 0] * sa + loc[:, 1] * ca + c[1],
                             loc[:, 2] + c[2]], 1))
    return np.concatenate(out).astype(np.float32)


def main():
    from oracle import lidar_ops as LO
    assert LO.ref_cpu_available(), "build oracle/_ref first (make -C oracle)"
    out = {}
    for seed in (0, 1):
        boxes = synth_boxes(seed)
        pts = synth_box_points(seed, boxes)
        got = LO.ref_points_in_boxes_cpu(pts, boxes)
        out[f"bits_{seed}"] = np.packbits(got.astype(np.uint8), axis=1)
        out[f"count_{seed}"] = np.array([got.sum()], np.int64)
        # the Python wrapper's view (roiaware_pool3d_utils.py:9-25): boxes enlarged by 0.2 m
        big = boxes.copy()
        big[:, 3:6] += 0.2
        out[f"bits_big_{seed}"] = np.packbits(LO.ref_points_in_boxes_cpu(pts, big).astype(np.uint8), axis=1)
    np.savez_compressed(os.path.join(HERE, "roiaware.npz"), **out)
    print({k: v.shape for k, v in out.items()}, {k: int(v[0]) for k, v in out.items() if k.startswith("count")})


if __name__ == "__main__":
    main()
