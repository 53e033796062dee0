"""Golden vectors of the temporal (autoregressive 4D) glue from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_temporal.py          -> tests/golden/temporal.npz, tests/golden/boxes2d.npz

Runs, on seeded synthetic scenes, the reference's own
  * lidargen/dataset/transforms_3d/common.py:99-216   convert_boxes_to_2d (float32 AND float64 boxes, seam-straddling boxes)
  * lidargen/dataset/nuscenes_dataset.py:375-421      NuscDataset.pre_process (tasks layout_cond / autoregressive_generation)
  * lidargen/dataset/custom_dataset.py:57-89          CustomDataset.__getitem__
  * tools/vis_tools/utils/pipe_related.py:28-95,243-288  get_temporal_boxes_3d / get_next_frame_points / delete_fg_points
  * tools/vis_tools/utils/common.py:59-222            warp_lidar_future / warp_boxes_future / compute_inter_frame_transforms
Import recipe: oracle/ref_import.py stubs the package __init__ files; here the heavy leaf imports the dataset classes
pull in (clip, pyquaternion, the scene-graph assigner, PTv3) are replaced by empty stand-ins, and the compiled
`roiaware_pool3d_cuda` extension by the reference's own C++ built into oracle/_ref (oracle/Makefile).
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

from oracle import lidar_ops as LO  # noqa: E402
from oracle import ref_import as R  # noqa: E402

CLASS_NAMES = ('car', 'truck', 'construction_vehicle', 'bus', 'trailer', 'motorcycle', 'bicycle', 'pedestrian')
H, W = 32, 1024


def import_reference():
    R.setup()
    ref = R.REF_ROOT

    def stub(name, **attrs):
        m = types.ModuleType(name)
        for k, v in attrs.items():
            setattr(m, k, v)
        sys.modules[name] = m
        return m

    class _Dummy:
        def __init__(self, *a, **k):
            pass

    stub("clip")
    stub("pyquaternion", Quaternion=_Dummy)
    for name, rel in [("lidargen.dataset.transforms_3d.scene_graph", "lidargen/dataset/transforms_3d/scene_graph"),
                      ("lidargen.dataset.augmentor", "lidargen/dataset/augmentor"),
                      ("lidargen.ops", "lidargen/ops"), ("lidargen.ops.roiaware_pool3d", "lidargen/ops/roiaware_pool3d"),
                      ("lidargen.metrics", "lidargen/metrics"), ("lidargen.metrics.models", "lidargen/metrics/models"),
                      ("lidargen.metrics.models.ptv3", "lidargen/metrics/models/ptv3"),
                      ("tools", "tools"), ("tools.vis_tools", "tools/vis_tools"), ("tools.vis_tools.utils", "tools/vis_tools/utils")]:
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(ref, rel)]
        sys.modules[name] = m
    stub("lidargen.dataset.transforms_3d.scene_graph.scene_graph", SceneGraphAssigner=_Dummy)
    stub("lidargen.dataset.augmentor.data_augmentor", DataAugmentor=_Dummy)
    stub("lidargen.metrics.models.ptv3.model", PTv3=_Dummy)

    def points_in_boxes_cpu(boxes_t, pts_t, out_t):     # signature of the compiled extension (roiaware_pool3d.cpp:144)
        out_t.copy_(torch.from_numpy(LO.ref_points_in_boxes_cpu(pts_t.numpy(), boxes_t.numpy())))
        return 1

    stub("lidargen.ops.roiaware_pool3d.roiaware_pool3d_cuda", points_in_boxes_cpu=points_in_boxes_cpu)
    from lidargen.dataset import custom_dataset
    from lidargen.dataset.transforms_3d import common as tcommon
    from tools.vis_tools.utils import common as vcommon
    from tools.vis_tools.utils import pipe_related
    return custom_dataset, tcommon, vcommon, pipe_related


def synth_scene(seed: int, n_obj: int = 12, dtype=np.float32):
    """boxes (ego row first) around the ego vehicle + names + a point cloud: ground plane, walls, and points ON the boxes"""
    rs = np.random.RandomState(seed)
    size = {'car': (4.67, 1.95, 1.74), 'truck': (7.12, 2.54, 2.90), 'construction_vehicle': (6.58, 2.75, 3.22),
            'bus': (11.23, 2.94, 3.49), 'trailer': (12.02, 2.91, 3.84), 'motorcycle': (2.07, 0.77, 1.44),
            'bicycle': (1.73, 0.62, 1.32), 'pedestrian': (0.77, 0.69, 1.78)}
    names = ['ego'] + [CLASS_NAMES[i] for i in rs.randint(0, 8, n_obj)]
    boxes = np.zeros((n_obj + 1, 7), dtype)
    for i in range(1, n_obj + 1):
        r = rs.uniform(5, 40)
        a = rs.uniform(-np.pi, np.pi) if i > 2 else np.pi - 0.02 * i * (-1) ** i        # two boxes at the azimuth seam
        boxes[i, :3] = [r * np.cos(a), r * np.sin(a), rs.uniform(-1.5, -0.5)]
        boxes[i, 3:6] = size[names[i]]
        boxes[i, 6] = rs.uniform(-np.pi, np.pi)
    # sweep: 32 rings x 1085 azimuths hitting the ground plane z = -1.8 or a wall at U(10, 60) m
    el = np.deg2rad(np.linspace(-30, 10, 32, endpoint=False) + 0.6)[:, None]
    az = np.linspace(-np.pi, np.pi, 1085, endpoint=False)[None, :]
    d_ground = np.where(np.sin(el) < -0.02, -1.8 / np.minimum(np.sin(el), -0.02), 1e9) + 0 * az
    d_wall = rs.uniform(10, 60, (1, 1085)) + 0 * el
    d = np.minimum(np.minimum(d_ground, d_wall), 80.0) * rs.uniform(0.98, 1.02, (32, 1085))
    pts = np.stack([d * np.cos(el) * np.cos(az), d * np.cos(el) * np.sin(az), d * np.sin(el), rs.uniform(0, 255, d.shape)], -1)
    pts = pts.reshape(-1, 4)
    on_boxes = []
    for b in boxes[1:].astype(np.float64):
        loc = rs.uniform(-0.5, 0.5, (300, 3)) * b[3:6]
        c, s = np.cos(b[6]), np.sin(b[6])
        on_boxes.append(np.stack([loc[:, 0] * c - loc[:, 1] * s + b[0], loc[:, 0] * s + loc[:, 1] * c + b[1], loc[:, 2] + b[2],
                                  rs.uniform(0, 255, 300)], 1))
    pts = np.concatenate([pts] + on_boxes).astype(np.float32)
    trajs = np.zeros((n_obj + 1, 6, 2), dtype)           # per-step (dx, dy) of ego (row 0) and objects, 6 future steps
    trajs[0, :, 0] = 0.5
    trajs[1:] = rs.uniform(-0.4, 0.4, (n_obj, 6, 2))
    trajs[3] = 0.0                                        # a parked object
    return boxes, names, pts, trajs


def main():
    custom_dataset, tcommon, vcommon, pipe = import_reference()
    out, out_b = {}, {}
    # ---- convert_boxes_to_2d: float64 boxes (previous fixture layout) + float32 boxes + seam cases ----
    for case in range(6):
        dtype = np.float64 if case < 3 else np.float32
        boxes, names, _, _ = synth_scene(100 + case, dtype=dtype)
        cls = np.array([(['ego'] + list(CLASS_NAMES)).index(n) for n in names], dtype=np.int32)
        b8 = np.concatenate((boxes, cls.reshape(-1, 1).astype(np.float32)), axis=1)
        b2, mask, w = tcommon.convert_boxes_to_2d(b8.copy(), H=H, W=W, fov_up=10.0, fov_down=-30.0)
        out_b[f"boxes_{case}"], out_b[f"b2_{case}"], out_b[f"mask_{case}"], out_b[f"w_{case}"] = b8, b2, mask, w
    np.savez_compressed(os.path.join(HERE, "boxes2d.npz"), **out_b)

    # ---- the dataset item (pre_process) and the temporal glue ----
    full = {}                                        # FULL=path.npz: every large array in full (debugging aid, not committed)

    def big(key, arr):
        """large outputs travel as SHA-256 of the bytes + shape + every 41st row (order-sensitive: the device path must
        produce the reference's row order)"""
        import hashlib
        arr = np.ascontiguousarray(arr)
        full[key] = arr
        out[key + "_sha"] = np.frombuffer(hashlib.sha256(arr.tobytes()).digest(), dtype=np.uint8)
        out[key + "_shape"] = np.array(arr.shape)
        out[key + "_rows"] = arr[::41].copy()

    for scene in range(1):
        boxes, names, pts, trajs = synth_scene(200 + scene)
        k = f"s{scene}_"
        out[k + "boxes"], out[k + "names"], out[k + "points"], out[k + "trajs"] = boxes, np.array(names), pts, trajs
        # first frame item: task layout_cond (what the box-layout sampler is fed)
        ds = custom_dataset.CustomDataset([dict(points=pts.copy(), gt_boxes=boxes.copy(), gt_names=list(names))])
        item = ds.__getitem__(0)
        for key in ("xyz", "reflectance", "depth", "mask", "scaled_gt_boxes", "fg_encoding_box", "gt_boxes_2d", "is_valid_obj",
                    "condition_mask", "scene_loss_weight_map", "gt_boxes"):
            out[k + "item_" + key] = np.asarray(item[key])
        first = dict(gt_fut_trajs=trajs.copy(), xyz=item["xyz"].copy(), reflectance=item["reflectance"].copy(),
                     gt_boxes=boxes.copy(), gt_names=list(names), condition_mask=item["condition_mask"].copy())
        bg, fut_bg, cur_boxes, fut_boxes, Ts, objs, inten = pipe.get_temporal_boxes_3d(first, M=None)
        out[k + "fut_boxes"], out[k + "Ts"] = fut_boxes, Ts
        big(k + "bg", bg)
        big(k + "fut_bg_0", fut_bg[0])
        big(k + "fut_bg_2", fut_bg[2])
        out[k + "obj_counts"] = np.array([o.shape[0] for o in objs])
        out[k + "obj_points"] = np.concatenate(objs)
        out[k + "obj_intensity"] = np.concatenate(inten)
        cur = bg
        for t in range(2):
            nxt = pipe.get_next_frame_points(cur, objs, inten, fut_boxes[:, t], list(names), Ts[t])
            big(k + f"next_{t}", nxt)
            gt = np.concatenate([np.zeros((1, 7), np.float32), fut_boxes[:, t]], axis=0)
            ds_t = custom_dataset.CustomDataset([dict(points=nxt.copy(), gt_boxes=gt.copy(), gt_names=list(names))])
            setattr(ds_t, "task", "autoregressive_generation")
            it = ds_t.__getitem__(0)
            out[k + f"ar_cond_{t}"] = it["autoregressive_cond"]
            out[k + f"ar_condition_mask_{t}"] = it["condition_mask"]
            out[k + f"ar_scaled_gt_boxes_{t}"] = it["scaled_gt_boxes"]
            out[k + f"ar_gt_boxes_2d_{t}"] = it["gt_boxes_2d"]
            # stand-in for the generated frame: the re-projected input, pushed 0.3 % outwards (an exact copy would put two
            # points of EQUAL depth into most pixels of the next projection, and the winner of a tie is whatever order the
            # reference's unstable np.argsort leaves -- not a property worth pinning; a sampled frame has no such ties)
            ds_g = custom_dataset.CustomDataset([dict(points=nxt.copy(), gt_boxes=gt.copy(), gt_names=list(names))])
            g = ds_g.__getitem__(0)
            gen = np.concatenate([g["xyz"] * np.float32(1.003), g["reflectance"]], 0).reshape(4, -1).T
            combined = np.concatenate([fut_bg[t], gen], axis=0)
            cur = pipe.delete_fg_points(combined, gt[1:, :7].copy())
            big(k + f"bg_after_{t}", cur)
    np.savez_compressed(os.path.join(HERE, "temporal.npz"), **out)
    if os.environ.get("FULL"):
        np.savez(os.environ["FULL"], **full)
    print({k: (v.shape, str(v.dtype)) for k, v in out.items()})


if __name__ == "__main__":
    main()
