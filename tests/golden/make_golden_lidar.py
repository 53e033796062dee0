"""Golden vectors for the point-cloud ops from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_lidar.py

load_points_as_images (lidargen/dataset/transforms_3d/common.py:26-91) is run on seeded synthetic nuScenes-shaped
sweeps; we store, per frame, the per-point bins and the index of the point that won
each pixel (which, together with the seeded points, determines the whole [H,W,6] image).
"""
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
warnings.filterwarnings("ignore")


def synth_sweep(seed: int, M: int = 34720):
    """32 rings x 1085 azimuth steps on the 10/-30 deg fov, jittered; depth = min(ground plane, U(2,60))."""
    rs = np.random.RandomState(seed)
    rings, az_n = 32, M // 32
    elev = np.deg2rad(np.linspace(-29.4, 9.4, rings))[:, None] + rs.normal(0, np.deg2rad(0.12), (rings, az_n))
    azim = np.linspace(-np.pi, np.pi, az_n, endpoint=False)[None, :] + rs.normal(0, 1e-3, (rings, az_n))
    d = rs.uniform(2.0, 60.0, (rings, az_n))
    ground = np.where(elev < -0.02, 1.8 / np.maximum(np.sin(-elev), 1e-3), 1e9)
    d = np.minimum(np.minimum(d, ground), 95.0)
    d[rs.rand(rings, az_n) < 0.01] = rs.uniform(0.3, 1.4)          # a few returns closer than min_depth
    x = d * np.cos(elev) * np.cos(azim)
    y = d * np.cos(elev) * np.sin(azim)
    z = d * np.sin(elev)
    inten = rs.uniform(0, 255, (rings, az_n))
    pts = np.stack([x, y, z, inten], -1).reshape(-1, 4).astype(np.float32)
    return pts[rs.permutation(len(pts))]


def main():
    from oracle import ref_import as R
    common = R.transforms_common()
    out = {}
    for seed in (0, 1, 2):
        pts = synth_sweep(seed)
        img = common.load_points_as_images(points=pts.copy(), scan_unfolding=False, H=32, W=1024, min_depth=1.45,
                                           max_depth=80.0, fov_up=10.0, fov_down=-30.0)
        # recover per-point bins exactly as the reference computes them (same expressions, common.py:72-85)
        x, y, z = pts[:, [0]], pts[:, [1]], pts[:, [2]]
        depth = np.linalg.norm(pts[:, :3], ord=2, axis=1, keepdims=True)
        h_up, h_down = np.deg2rad(10.0), np.deg2rad(-30.0)
        elevation = np.arcsin(z / (depth + 1e-6)) + abs(h_down)
        grid_h = np.floor((1 - elevation / (h_up - h_down)) * 32).clip(0, 31).astype(np.int32)
        azimuth = -np.arctan2(y, x)
        grid_w = np.floor(((azimuth / np.pi + 1) / 2 % 1) * 1024).clip(0, 1023).astype(np.int32)
        grid = np.concatenate([grid_h, grid_w], 1)
        # winner index per pixel from the reference image (match by exact xyz + intensity)
        key = {}
        for i, p in enumerate(pts):
            key[(p[0], p[1], p[2], p[3])] = i
        win = -np.ones((32, 1024), np.int32)
        nz = np.argwhere(img[..., 4] > 0)
        for h, w in nz:
            win[h, w] = key[tuple(img[h, w, :4])]
        out[f"grid_{seed}"] = grid.astype(np.int16)
        out[f"win_{seed}"] = win
        out[f"imgsum_{seed}"] = np.array([img.astype(np.float64).sum(), (img[..., 5] > 0).sum()])
        print(seed, "occupied", (win >= 0).sum(), "mask", int(img[..., 5].sum()))
    np.savez_compressed(os.path.join(HERE, "projection.npz"), **out)


if __name__ == "__main__":
    main()
