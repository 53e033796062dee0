"""Layout-conditioned clip generation: first frame from the box-layout denoiser, then the autoregressive 4D rollout
(reference: tools/evaluation/sample_and_save_temporal.py:27-333 without the file output; BASELINE configs[2] = a 5-frame
clip, configs[4] = a 20-frame rollout).

Per clip the reference alternates between the GPU (two `sample()` loops) and NumPy on the host (dataset items, box
rasterisation, ego-motion warp, object pasting, re-projection, 3x points_in_boxes_cpu per frame and sample).  Here the
whole frame loop stays on the device: the box quantities of EVERY frame are known before the loop starts (trajectories are
inputs), so their 2-D boxes / condition masks are rasterised for all frames and samples in two launches; the point-set
glue is lidarcrafter_b200.rollout (fixed-capacity buffers, device-side counts, no host synchronisation per frame); the
denoiser plans / CUDA graphs of the two diffusion models are replayed per step as in `sample()`.
"""
from __future__ import annotations

import numpy as np
import torch

from . import layout_ops as LO
from . import ops, rollout


def interp_trajs_numpy(trajs: np.ndarray, M: int) -> np.ndarray:
    """pipe_related.py:230-241: (K,N,2) -> (K,M,2) by linear interpolation over [0,1]"""
    K, N, _ = trajs.shape
    t0, t1 = np.linspace(0.0, 1.0, N), np.linspace(0.0, 1.0, M)
    out = np.zeros((K, M, 2), dtype=trajs.dtype)
    for k in range(K):
        out[k, :, 0] = np.interp(t1, t0, trajs[k, :, 0])
        out[k, :, 1] = np.interp(t1, t0, trajs[k, :, 1])
    return out


def _per_step(trajs: np.ndarray, M: int | None = None) -> np.ndarray:
    """insert a zero step, accumulate, (interpolate to M poses,) difference: the per-step offsets the reference re-derives
    twice (sample_and_save_temporal.py:262-266 with interpolation, then pipe_related.py:29-34 without)"""
    acc = np.cumsum(np.insert(trajs, 0, 0, axis=1), axis=1)
    if M is not None:
        acc = interp_trajs_numpy(acc, M=M)
    return acc[:, 1:] - acc[:, :-1]


def future_trajectories(gt_fut_trajs: np.ndarray, traj_length: int | None = None, resample: bool = True):
    """per-step offsets (K+1,T,2), ego first -> (ego cumulative (T',2), objects cumulative (K,T',2)).  ``resample`` = the
    sampling script's own pass in front of get_temporal_boxes_3d (which repeats the same arithmetic without interpolation)"""
    step = np.asarray(gt_fut_trajs)
    if resample:
        step = _per_step(step, traj_length)
    step = _per_step(step)
    return np.cumsum(step[0], axis=0), np.cumsum(step[1:], axis=1)


class ClipState:
    """per-sample device state of a rollout (what the reference keeps in six Python lists, :248-281)"""
    __slots__ = ("bg0", "bg", "obj_points", "obj_intensity", "obj_box", "fut_boxes", "Ts", "ego_xy", "names")


class TemporalSampler:
    """first frame (`ddpm`, box-layout model) + autoregressive frames (`auto_ddpm`, in_channels 2 + 10 + 1)."""

    def __init__(self, ddpm, auto_ddpm, lidar_utils, resolution=(32, 1024), min_depth: float = 1.45, max_depth: float = 80.0,
                 fov_up: float = 10.0, fov_down: float = -30.0, class_names=LO.CLASS_NAMES):
        self.ddpm, self.auto_ddpm, self.lidar_utils = ddpm, auto_ddpm, lidar_utils
        self.H, self.W = resolution
        self.geom = dict(H=self.H, W=self.W, min_depth=min_depth, max_depth=max_depth, fov_up=fov_up, fov_down=fov_down)
        self.class_names = class_names

    @property
    def device(self):
        return self.ddpm.device

    # ---- conditioning ----------------------------------------------------------------------------------------------
    def box_batch(self, boxes_list, names_list, dtype=np.float32) -> dict:
        """the collated dataset items of B samples (CustomDataset.__getitem__ -> pre_process -> collate_fn,
        custom_dataset.py:57-89, base_dataset.py:38-75) for given [N+1,7] boxes (ego row first): scaled_gt_boxes [B,13,9],
        gt_boxes_2d [B,13,4], is_valid_obj [B,13], condition_mask [B,2,H,W] on the device"""
        dev = self.device
        B = len(boxes_list)
        n_max = max(b.shape[0] for b in boxes_list)
        boxes8 = np.zeros((B, n_max, 8), dtype)
        scaled, valid = np.zeros((B, 13, 9)), np.zeros((B, 13))
        for i, (b, names) in enumerate(zip(boxes_list, names_list)):
            b8 = np.concatenate((np.asarray(b, dtype), LO.class_ids(names, self.class_names).reshape(-1, 1).astype(np.float32)),
                                axis=1).astype(dtype)
            boxes8[i, :b8.shape[0]] = b8
            n = min(13, b8.shape[0] - 1)
            scaled[i, :n] = LO.scale_boxes_3d(b8.copy())[1:1 + n]
            valid[i, :n] = 1
        # rows beyond a sample's box count are all-zero boxes: they project to an empty rectangle (x1 == x2) and draw nothing
        b2, mask, _ = LO.convert_boxes_to_2d(torch.from_numpy(boxes8), **self.geom)
        b2p = torch.zeros(B, 13, 4, dtype=torch.float64, device=dev)
        for i, b in enumerate(boxes_list):
            n = min(13, b.shape[0] - 1)
            b2p[i, :n] = b2[i, 1:1 + n]
        return dict(scaled_gt_boxes=torch.from_numpy(scaled).float().to(dev), gt_boxes_2d=b2p.float(),
                    is_valid_obj=torch.from_numpy(valid).float().to(dev), condition_mask=mask)

    def prepare_batch(self, batch: dict, autoregressive_cond: torch.Tensor | None = None) -> dict:
        """sample_and_save_temporal.py:151-177: condition_mask -> concat_cond (one-hot classes + normalised depth),
        autoregressive_cond (depth, reflectance) -> normalised depth channel"""
        out = dict(batch)
        out["concat_cond"] = LO.preprocess_condition_mask(batch["condition_mask"], self.lidar_utils)
        if autoregressive_cond is not None:
            out["autoregressive_cond"] = LO.preprocess_autoregressive_cond(autoregressive_cond, self.lidar_utils, (self.H, self.W))
        return out

    def postprocess(self, sample: torch.Tensor) -> torch.Tensor:
        """:194-199: [B,2,H,W] in [-1,1] -> [B,5,H,W] (metric depth, x, y, z, reflectance in [0,1])"""
        depth, xyz = self.lidar_utils.to_xyz_from_normalized(sample[:, [0]].contiguous())
        return torch.cat([depth, xyz, self.lidar_utils.denormalize(sample[:, [1]])], dim=1)

    # ---- clip set-up (pipe_related.py:28-95) ------------------------------------------------------------------------
    def start_clip(self, frame: torch.Tensor, boxes: np.ndarray, names, gt_fut_trajs: np.ndarray, condition_mask: torch.Tensor,
                   traj_length: int | None = None, resample: bool = True) -> ClipState:
        """frame [5,H,W] (postprocess output of sample 0), boxes [N+1,7] -> per-object canonical points, background,
        future boxes / transforms (get_temporal_boxes_3d)"""
        st = ClipState()
        ego_xy, obj_xy = future_trajectories(gt_fut_trajs, traj_length, resample)
        cur_boxes = np.array(np.asarray(boxes)[1:, :7], copy=True)
        xyz, inten = frame[1:4], frame[4:5] * 255
        rows, _ = rollout.image_points(xyz, inten)
        st.obj_points, st.obj_intensity, st.obj_box = rollout.extract_object_points(rows, rollout.remove_ego_points(rows), cur_boxes)
        rows_bg, valid_bg = rollout.image_points(xyz, inten, ~(condition_mask[0] > 0))
        st.bg0 = rollout.compact(rows_bg, valid_bg)
        st.bg = st.bg0
        if cur_boxes.dtype == np.float32:
            # points_in_boxes_cpu enlarges a float32 box array IN PLACE by 0.2 m (roiaware_pool3d_utils.py:21 through
            # check_numpy_to_torch's shared memory): every future box of the reference carries that margin
            cur_boxes[:, 3:6] += np.float32(0.2)
        st.fut_boxes = rollout.warp_boxes_future(cur_boxes, obj_xy, ego_xy, 0.0)          # (K, T, 7), dtype of the boxes
        st.Ts = rollout.compute_inter_frame_transforms(ego_xy, 0.0)
        st.ego_xy, st.names = ego_xy, list(names)
        return st

    # ---- the clip ------------------------------------------------------------------------------------------------------
    @torch.inference_mode()
    def generate(self, scenes: list, num_frames: int, num_steps: int, mode: str = "ddim", temporal_mode: str = "ddpm",
                 rng=None, traj_length: int | None = None, progress: bool = False) -> torch.Tensor:
        """scenes: B dicts {gt_boxes [N+1,7] (ego row first), gt_names [N+1], gt_fut_trajs [N+1,T,2]} ->
        [B, num_frames, 5, H, W] (depth, x, y, z, reflectance per frame).  Mirrors sample_and_save_temporal.py:203-333
        (the reference hard-codes 15 future frames; here num_frames - 1 <= T)."""
        B = len(scenes)
        H, W = self.H, self.W
        boxes0 = [np.asarray(s["gt_boxes"]) for s in scenes]
        names = [list(s["gt_names"]) for s in scenes]
        # ---- frame 0: box-layout model ----
        batch = self.prepare_batch(self.box_batch(boxes0, names, dtype=boxes0[0].dtype if boxes0[0].dtype in (np.float32, np.float64) else np.float32))
        x = self.ddpm.sample(batch_dict=batch, batch_size=B, num_steps=num_steps, mode=mode, progress=progress, rng=rng).clamp(-1, 1)
        frames = [self.postprocess(x)]
        if num_frames == 1:
            return torch.stack(frames, dim=1)
        states = [self.start_clip(frames[0][b], boxes0[b], names[b], scenes[b]["gt_fut_trajs"], batch["condition_mask"][b],
                                  traj_length) for b in range(B)]
        T_fut = num_frames - 1
        assert all(st.fut_boxes.shape[1] >= T_fut for st in states), "trajectories shorter than the clip"
        # ---- box conditioning of every future frame, all samples, in two rasteriser launches per dtype flow ----
        ego32, ego64 = np.zeros((1, 7), np.float32), np.zeros((1, 7))
        fut32 = [np.concatenate([ego32, st.fut_boxes[:, t]], axis=0) for t in range(T_fut) for st in states]     # dataset items (:300-307)
        fut64 = [np.concatenate([ego64, st.fut_boxes[:, t]], axis=0) for t in range(T_fut) for st in states]     # refine step (pipe_related.py:251-255)
        names_t = [n for _ in range(T_fut) for n in names]
        cond_items = self.box_batch(fut32, names_t, dtype=np.result_type(np.float32, states[0].fut_boxes.dtype).type)
        refine_mask = self.box_batch(fut64, names_t, dtype=np.float64)["condition_mask"]
        for t in range(T_fut):
            sl = slice(t * B, (t + 1) * B)
            nxt = [rollout.get_next_frame_points(st.bg, st.obj_points, st.obj_intensity, st.obj_box, st.fut_boxes[:, t], st.Ts[t],
                                                 refine_mask[t * B + b], **self.geom) for b, st in enumerate(states)]
            # CustomDataset(task='autoregressive_generation').__getitem__ (custom_dataset.py:59-80): re-project the assembled cloud
            cap = max(p.buf.shape[0] for p in nxt)
            pts = torch.zeros(B, cap, 4, device=self.device)
            for b, p in enumerate(nxt):
                pts[b, :p.buf.shape[0]] = p.buf
            img = ops.load_points_as_images(points=pts, npts=torch.cat([p.n for p in nxt]), **self.geom)      # [B,H,W,6]
            img = (img * img[..., 5:6]).permute(0, 3, 1, 2)
            ar = torch.cat([img[:, 4:5], rollout.div255(img[:, 3:4])], dim=1)                                            # depth, reflectance
            bd = self.prepare_batch({k: v[sl] for k, v in cond_items.items()}, autoregressive_cond=ar)
            x = self.auto_ddpm.sample(batch_dict=bd, batch_size=B, num_steps=num_steps, mode=temporal_mode, progress=progress,
                                      rng=rng).clamp(-1, 1)
            fr = self.postprocess(x)
            frames.append(fr)
            for b, st in enumerate(states):
                gen = fr[b, 1:5].reshape(4, -1).T                                                               # x, y, z, reflectance (:321)
                fut_bg = rollout.warp_lidar_future(st.bg0.buf, st.ego_xy, t)
                n0 = st.bg0.buf.shape[0]
                comb = torch.cat([fut_bg, gen], dim=0)
                valid = torch.cat([st.bg0.valid, torch.ones(gen.shape[0], dtype=torch.bool, device=gen.device)])
                st.bg = rollout.delete_fg_points(rollout.compact(comb, valid), fut32[t * B + b][1:, :7])
                del n0
        return torch.stack(frames, dim=1)
