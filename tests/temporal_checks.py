"""Checks of the temporal (autoregressive 4D) glue against tests/golden/temporal.npz = outputs of the reference's OWN
functions (tests/golden/make_golden_temporal.py).  Shared by the CPU suite (C-ABI emulator, whose point-cloud entries are
the C oracle) and the GPU suite (the real library): the caller decides which library is active.

Large reference arrays travel as SHA-256 + shape + every 41st row.  Index / selection results (which pixels, which points
survive, the 2-D boxes, the masks) are compared EXACTLY; float coordinates that went through a BLAS matmul in the
reference (object rotation, ego warp in float32) within 2e-6 relative."""
import hashlib
import os

import numpy as np
import torch

from lidarcrafter_b200 import layout_ops as LO
from lidarcrafter_b200 import ops, rollout, temporal

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "temporal.npz"))
K = "s0_"
GEOM = dict(H=32, W=1024, min_depth=1.45, max_depth=80.0, fov_up=10.0, fov_down=-30.0)


def sha(a: np.ndarray) -> np.ndarray:
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def same_big(key: str, got: np.ndarray, exact: bool, rtol: float = 2e-6, atol: float = 2e-5):
    assert tuple(G[key + "_shape"]) == got.shape, (key, tuple(G[key + "_shape"]), got.shape)
    if exact:
        assert np.array_equal(sha(got), G[key + "_sha"]), key
    else:
        assert np.allclose(got[::41], G[key + "_rows"], rtol=rtol, atol=atol), (key, np.abs(got[::41] - G[key + "_rows"]).max())


def same_set(key: str, got: np.ndarray, min_frac: float = 0.999, tol: float = 1e-4):
    """order-insensitive variant for the frames after the first: a pasted object point that lands one pixel aside (BLAS
    rounding of the reference's rotation) removes / adds a row and shifts everything behind it"""
    from scipy.spatial import cKDTree
    ref = G[key + "_rows"]
    n_ref = int(G[key + "_shape"][0])
    assert abs(got.shape[0] - n_ref) <= max(4, n_ref // 2000), (key, got.shape[0], n_ref)
    d, _ = cKDTree(got.astype(np.float64)).query(ref.astype(np.float64))
    assert (d < tol * (1 + np.linalg.norm(ref, axis=1))).mean() >= min_frac, (key, float((d < tol).mean()))


def check_layout_item():
    it = LO.layout_item(G[K + "boxes"], list(G[K + "names"]), **GEOM)
    for k in ("scaled_gt_boxes", "fg_encoding_box", "gt_boxes_2d", "is_valid_obj", "condition_mask", "gt_boxes"):
        assert np.array_equal(np.asarray(it[k]), G[K + "item_" + k]), k
    assert np.allclose(it["scene_loss_weight_map"], G[K + "item_scene_loss_weight_map"], rtol=1e-6)


def check_clip_glue(dev: str):
    """start_clip (get_temporal_boxes_3d) -> two frames of get_next_frame_points / dataset item / delete_fg_points"""
    names = list(G[K + "names"])
    frame = torch.from_numpy(np.concatenate([G[K + "item_depth"], G[K + "item_xyz"], G[K + "item_reflectance"]], 0)).to(dev)
    cmask = torch.from_numpy(G[K + "item_condition_mask"]).to(dev)

    class _D:
        device = torch.device(dev)
    ts = temporal.TemporalSampler(_D(), None, None)
    st = ts.start_clip(frame, G[K + "boxes"], names, G[K + "trajs"], cmask, resample=False)
    same_big(K + "bg", st.bg0.numpy(), exact=True)
    assert np.allclose(st.fut_boxes, G[K + "fut_boxes"], rtol=1e-6, atol=1e-6) and st.fut_boxes.dtype == G[K + "fut_boxes"].dtype
    assert np.allclose(st.Ts, G[K + "Ts"], atol=1e-12)
    counts = np.bincount(st.obj_box.cpu().numpy(), minlength=len(names) - 1)
    assert np.array_equal(counts, G[K + "obj_counts"])
    assert np.allclose(st.obj_points.cpu().numpy(), G[K + "obj_points"], rtol=2e-6, atol=2e-6)
    assert np.array_equal(st.obj_intensity.cpu().numpy(), G[K + "obj_intensity"])
    same_big(K + "fut_bg_0", rollout.warp_lidar_future(st.bg0.buf, st.ego_xy, 0)[:int(st.bg0.n)].cpu().numpy(), exact=False)
    same_big(K + "fut_bg_2", rollout.warp_lidar_future(st.bg0.buf, st.ego_xy, 2)[:int(st.bg0.n)].cpu().numpy(), exact=False)

    fut = G[K + "fut_boxes"]                     # the reference's future boxes: identical inputs for the index-exact checks
    for t in range(2):
        gt64 = np.concatenate([np.zeros((1, 7)), fut[:, t]], axis=0)
        gt32 = np.concatenate([np.zeros((1, 7), np.float32), fut[:, t]], axis=0)
        refine_mask = ts.box_batch([gt64], [names], dtype=np.float64)["condition_mask"][0]
        nxt = rollout.get_next_frame_points(st.bg, st.obj_points, st.obj_intensity, st.obj_box, fut[:, t], G[K + "Ts"][t],
                                            refine_mask, **GEOM)
        got = nxt.numpy()
        if t == 0:
            same_big(K + f"next_{t}", got, exact=False)
            n_fg = st.obj_points.shape[0]
            # the refined background part (re-projection of float64 points, mask, distance filter) is index work: exact
            assert np.array_equal(got[:-n_fg:41], G[K + f"next_{t}_rows"][:len(got[:-n_fg:41])])
        else:
            same_set(K + f"next_{t}", got)
        # dataset item of the autoregressive task (custom_dataset.py:59-80 + pre_process)
        item = ts.box_batch([gt32], [names], dtype=np.float32)
        assert np.array_equal(item["condition_mask"][0].cpu().numpy(), G[K + f"ar_condition_mask_{t}"])
        assert np.array_equal(item["scaled_gt_boxes"][0].cpu().numpy(), G[K + f"ar_scaled_gt_boxes_{t}"].astype(np.float32))
        assert np.array_equal(item["gt_boxes_2d"][0].cpu().numpy(), G[K + f"ar_gt_boxes_2d_{t}"].astype(np.float32))
        img = ops.load_points_as_images(points=nxt.buf[None], npts=nxt.n, **GEOM)[0]
        img = (img * img[..., 5:6]).permute(2, 0, 1)
        ar = torch.cat([img[4:5], rollout.div255(img[3:4])], 0).cpu().numpy()
        ref_ar = G[K + f"ar_cond_{t}"]
        # pasted object points differ from the reference's BLAS rotation in the last bits: a handful may change pixel
        assert ((ar != 0) != (ref_ar != 0)).mean() < 2e-4
        assert (np.abs(ar - ref_ar) > 1e-4 * (1 + np.abs(ref_ar))).mean() < 1e-3
        # stand-in for the generated frame (as in the golden script): the re-projected cloud, pushed 0.3 % outwards
        gen = torch.cat([img[:3] * 1.003, rollout.div255(img[3:4])], 0).reshape(4, -1).T
        fut_bg = rollout.warp_lidar_future(st.bg0.buf, st.ego_xy, t)
        comb = rollout.compact(torch.cat([fut_bg, gen], 0),
                               torch.cat([st.bg0.valid, torch.ones(gen.shape[0], dtype=torch.bool, device=gen.device)]))
        st.bg = rollout.delete_fg_points(comb, gt32[1:, :7])
        same_set(K + f"bg_after_{t}", st.bg.numpy())
