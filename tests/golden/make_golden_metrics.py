"""Golden vectors for the evaluation-side point-cloud ops from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_metrics.py

lidargen/metrics/metric_utils.py (pcd2range, range2xyz, sparse_quantize, pcd2bev_sum, pcd2voxel_full, pcd2bev_bin,
bev_sample) is run on the seeded nuScenes-shaped sweeps of make_golden_lidar.synth_sweep.  Small outputs are stored
whole, large ones as SHA-256 digests of their bytes (integer / exactly reproducible outputs) plus a few statistics.
"""
import hashlib
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

NUSC = dict(size=[32, 1024], fov=[10, -30], depth_range=[1.0, 45.0])
KITTI = dict(size=[64, 1024], fov=[3, -25], depth_range=[1.0, 56.0], depth_scale=6)


def digest(a) -> np.ndarray:
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8).copy()


def small_cloud(seed: int, n: int = 6000):
    """clustered points so that many share a voxel (exercises first-occurrence / inverse bookkeeping)"""
    rs = np.random.RandomState(seed)
    c = rs.uniform(-25, 25, (60, 3)) * np.array([1, 1, 0.1])
    return (c[rs.randint(0, 60, n)] + rs.normal(0, 0.08, (n, 3))).astype(np.float32)


def main():
    from make_golden_lidar import synth_sweep
    from oracle import ref_import as R
    MU = R.metric_utils()
    out = {}
    for seed in (0, 1):
        sw = synth_sweep(seed)
        pts, inten = sw[:, :3].copy(), sw[:, 3].copy()
        r, f = MU.pcd2range(pts, remission=inten, **NUSC)
        out[f"range_{seed}"] = r
        out[f"feat_digest_{seed}"] = digest(f)
        out[f"range_kitti_digest_{seed}"] = digest(MU.pcd2range(pts, **KITTI)[0])
        lab = (np.arange(len(pts)) % 17).astype(np.int64)
        out[f"label_digest_{seed}"] = digest(MU.pcd2range(pts, labels=lab, **NUSC)[1])
        xyz = MU.range2xyz(r, log_scale=False, depth_scale=None, **NUSC)
        out[f"xyz_sub_{seed}"] = xyz[:, ::2, ::16].copy()
        out[f"xyz_abs_sum_{seed}"] = np.array([np.abs(xyz).sum()])
        pr = MU.preprocess_range(pts, **KITTI)      # (the nuScenes config has no depth_scale: range2xyz would raise)
        out[f"prep_range_digest_{seed}"] = digest(pr[0].astype(np.float32))
        out[f"prep_range_abs_sum_{seed}"] = np.array([np.abs(pr).sum()])
        # sparse_quantize on the raw coordinates (fp64 division inside) with a 5 cm voxel: all three outputs
        c, i, inv = MU.sparse_quantize(pts, VOX := 0.05, return_index=True, return_inverse=True)
        out[f"sq_n_{seed}"] = np.array([len(c)])
        out[f"sq_coords_digest_{seed}"] = digest(c.astype(np.int32))
        out[f"sq_index_digest_{seed}"] = digest(i.astype(np.int64))
        out[f"sq_inverse_digest_{seed}"] = digest(inv.astype(np.int64))
        out[f"hash_digest_{seed}"] = digest(MU.ravel_hash(np.floor(pts / np.array([VOX] * 3)).astype(np.int32)))
        # bev / volume grids
        clouds = [pts, pts[: 5000] + np.float32(0.3), small_cloud(seed)]
        bs = MU.pcd2bev_sum('32', clouds)[0]
        nz = np.flatnonzero(bs)
        out[f"bev_sum_digest_{seed}"] = digest(np.stack([nz, bs.ravel()[nz].astype(np.int64)]))
        out[f"bev_sum_total_{seed}"] = np.array([bs.sum(), bs.max()])
        vf = MU.pcd2voxel_full('32', [pts])[0][0]
        out[f"voxel_full_digest_{seed}"] = digest(np.flatnonzero(vf))
        out[f"voxel_full_total_{seed}"] = np.array([vf.sum()])
        bb = MU.pcd2bev_bin('32', [pts, small_cloud(seed)])[0]
        out[f"bev_bin_digest_{seed}"] = digest(np.concatenate(bb))
        out[f"bev_bin_n_{seed}"] = np.array([len(b) for b in bb])
        sm = MU.bev_sample('32', [pts, small_cloud(seed)])[0]
        out[f"bev_sample_digest_{seed}"] = digest(np.concatenate(sm))
    # one small case stored whole
    sc = small_cloud(5)
    c, i, inv = MU.sparse_quantize(sc, (0.2, 0.2, 0.1), return_index=True, return_inverse=True)
    out["small_coords"], out["small_index"], out["small_inverse"] = c.astype(np.int32), i.astype(np.int32), inv.astype(np.int32)
    c2, i2 = MU.sparse_quantize(sc[:, :2].astype(np.float64), 0.25, return_index=True)
    out["small2d_coords"], out["small2d_index"] = c2.astype(np.int32), i2.astype(np.int32)
    # log-scale range2xyz (kitti config): exp2 in fp32
    rk = MU.pcd2range(synth_sweep(0)[:, :3], **KITTI)[0]
    logimg = (np.log2(np.maximum(rk, 0) + 1) / 6).astype(np.float32)
    xyzk = MU.range2xyz(logimg, **KITTI)
    out["xyz_log_sub"] = xyzk[:, ::4, ::16].copy()
    np.savez_compressed(os.path.join(HERE, "metrics.npz"), **out)
    print("wrote metrics.npz:", {k: (v.shape if v.size > 4 else v.tolist()) for k, v in out.items() if "digest" not in k})


if __name__ == "__main__":
    main()
