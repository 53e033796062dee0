#!/usr/bin/env python
"""bench.py -- denoiser sample-steps/s on the nuScenes 32x1024 range image (BASELINE.json metric).

A "step" is one pass of the hot path over one batch: one EfficientUNet forward + one DDIM update for a
batch of 8 frames per GPU (BASELINE.json configs[1]: single-frame diffusion, 50-step DDIM, batch 8).
`value` = samples * steps / time over all ranks (weak scaling: 8 frames per GPU, one NCCL all-gather of
the final frames).  Prints ONE JSON line (rank 0).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--precision fp16x3|fp16]
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

RES = (32, 1024)
NRES = (3, 3, 3, 3)
BATCH_PER_GPU = 8
METRIC = "denoiser_sample_steps_per_sec"
UNIT = "sample-steps/s"
# conv precision of the measured arm (lidarcrafter_b200/engine.py PRECISION_PARTS); all three meet or are reported
# against the 1e-3 relative tolerance of north_star: fp16x3 ~2e-6, fp16f8 ~5e-5, fp16 ~1.7e-3 (outside -> never default)
DEFAULT_PRECISION = "fp16f8"
DTYPE_NOTE = {"fp16x3": "f32 (fp16x3 split tensor-core MMAs, fp32 accumulate)",
              "fp16f8": "f32 (fp16 MMA + e4m3 correction MMA per product, fp32 accumulate; ~5e-5 rel. vs fp32)",
              "fp16": "f16 operands, f32 accumulate"}


# dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant conv shape (32x1024, 64 -> 64, 3x3) from the committed
# ncu --set full captures: mean of the launches without / with the residual read (profiles/r02_ncu_summary.md; the part of the
# output that is still in L2 when the kernel ends is not in dram__bytes_write)
NCU_TRAFFIC_32x1024_C64 = 1.42e8


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tf": d["bf16_tflops"], "tf_sustained": d.get("bf16_tflops_sustained"),
                "src": "measured (MEASURED_PEAKS.json)"}
    return {"hbm_gbs": 6650.0, "tf": 1590.0, "tf_sustained": 1400.0, "src": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "100", "-i", str(self.idx)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def random_weights(sd: dict, seed: int = 0) -> dict:
    """Random-init weights of the named architecture (there are no checkpoints offline).  The reference zero-inits the
    last conv of every branch, which would make the timed network degenerate, so every floating parameter is redrawn:
    GroupNorm gamma ~ N(1, 0.1), beta ~ N(0, 0.1), biases ~ N(0, 0.02), weights ~ N(0, 1/fan_in); one generator per key
    (same recipe as the test oracle uses, restated here so that the measured arm never imports oracle/)."""
    import math
    import zlib
    out = {}
    for k, v in sd.items():
        leaf = k.split(".")[-1]
        if leaf in ("coords", "scale", "kernel", "freqs", "phase", "_dummy") or not v.is_floating_point():
            out[k] = v.clone()
            continue
        g = torch.Generator().manual_seed((zlib.crc32(k.encode()) + 7919 * seed) % (2 ** 31))
        if leaf == "weight" and v.dim() == 1:
            out[k] = 1 + 0.1 * torch.randn(v.shape, generator=g)
        elif "norm" in k and leaf == "bias" and "proj" not in k:
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif leaf in ("bias", "in_proj_bias") or v.dim() == 1:
            out[k] = 0.02 * torch.randn(v.shape, generator=g)
        else:
            out[k] = torch.randn(v.shape, generator=g) / math.sqrt(v[0].numel())
    return out


def build_model(device, precision):
    import lidarcrafter_b200 as L
    m = L.EfficientUNet(in_channels=2, resolution=RES, base_channels=64, channel_multiplier=(1, 2, 4, 8),
                        num_residual_blocks=NRES, gn_num_groups=8, gn_eps=1e-6, attn_num_heads=8,
                        coords_encoding="fourier_features", ring=True)
    m.coords = L.get_linear_ray_angles(RES[0], RES[1], 10, -30)
    m.load_state_dict(random_weights(m.state_dict(), seed=0))
    m.precision = precision
    m = m.to(device).eval()
    ddpm = L.ContinuousTimeGaussianDiffusion(m, prediction_type="eps", noise_schedule="cosine").to(device)
    return m, ddpm


# LayoutUnetV1 / LayoutTransformerEncoder of the nuScenes box-layout configs (lidargen/utils/model/option_nusc_box_layout_v3.py
# ModelConfig.params / ConditionModelConfig.params; the 'nuscenes-auto-reg-v2' variant only differs by in_channels = 13)
LAYOUT_UNET = {'image_size': 32, 'use_fp16': False, 'use_scale_shift_norm': True, 'out_channels': 2, 'model_channels': 64,
               'encoder_channels': 64, 'num_head_channels': 32, 'num_heads': -1, 'num_heads_upsample': -1,
               'num_res_blocks': 2, 'num_attention_blocks': 1, 'resblock_updown': True, 'attention_ds': [4, 8],
               'channel_mult': [1, 2, 4, 8], 'dropout': 0.1, 'use_checkpoint': False,
               'use_positional_embedding_for_attention': True, 'attention_block_type': 'ObjectAwareCrossAttention'}
LAYOUT_ENC = {'feature_map_size': [32, 1024], 'used_condition_types': ['obj_class', 'obj_bbox', 'is_valid_obj'],
              'layout_length': 13, 'num_classes_for_layout_object': 9, 'mask_size_for_layout_object': 32, 'hidden_dim': 64,
              'output_dim': 256, 'num_layers': 6, 'num_heads': 4, 'use_final_ln': True, 'use_positional_embedding': False,
              'not_use_layout_fusion_module': False, 'resolution_to_attention': [4, 8], 'use_key_padding_mask': False,
              'out_channels': 10}
CLASS_SIZE = {'car': (4.67, 1.95, 1.74), 'truck': (7.12, 2.54, 2.90), 'construction_vehicle': (6.58, 2.75, 3.22),
              'bus': (11.23, 2.94, 3.49), 'trailer': (12.02, 2.91, 3.84), 'motorcycle': (2.07, 0.77, 1.44),
              'bicycle': (1.73, 0.62, 1.32), 'pedestrian': (0.77, 0.69, 1.78)}


def build_temporal(device, precision):
    """first-frame box-layout model + autoregressive model + the clip driver (tools/evaluation/sample_and_save_temporal.py:51-57)"""
    import lidarcrafter_b200 as L
    from lidarcrafter_b200.temporal import TemporalSampler
    pair = []
    for cin, seed in ((12, 0), (13, 2)):
        m = L.unets.__all__["layout_unet_v1"](in_channels=cin, resolution=RES, **LAYOUT_UNET)
        m.coords = L.get_linear_ray_angles(RES[0], RES[1], 10, -30)
        m.load_state_dict(random_weights(m.state_dict(), seed=seed))
        if hasattr(m, "precision"):
            m.precision = precision
        enc = L.unets.__all__["layout_encoder"](**dict(LAYOUT_ENC, out_channels=cin - 2))    # concat_cond 10 (+ 1 autoregressive depth)
        enc.load_state_dict(random_weights(enc.state_dict(), seed=seed + 1))
        pair.append(L.CondContinuousTimeGaussianDiffusion(m.eval(), enc.eval(), prediction_type="eps", noise_schedule="cosine",
                                                          cond_mode="concat").to(device))
    lu = L.LiDARUtility(RES, "log_depth", 1.45, 80.0, L.get_linear_ray_angles(RES[0], RES[1], 10, -30)).to(device)
    return TemporalSampler(pair[0], pair[1], lu, resolution=RES), pair


def synth_scenes(n: int, frames: int, seed: int = 0):
    """12 boxes per sample, centre (x, y) ~ U(-40, 40) outside |.| < 3, z ~ U(-2, 0), class-mean sizes,
    yaw ~ U(-pi, pi), class ~ U{1..8}; ego motion 0.5 m forward per frame, objects drift U(-0.4, 0.4) m per frame"""
    import numpy as np
    names_all = list(CLASS_SIZE)
    out = []
    for i in range(n):
        rs = np.random.RandomState(1000 * seed + i)
        names = ["ego"] + [names_all[k] for k in rs.randint(0, 8, 12)]
        boxes = np.zeros((13, 7), np.float32)
        for j in range(1, 13):
            while True:
                xy = rs.uniform(-40, 40, 2)
                if np.abs(xy).max() >= 3:
                    break
            boxes[j, :2], boxes[j, 2] = xy, rs.uniform(-2, 0)
            boxes[j, 3:6], boxes[j, 6] = CLASS_SIZE[names[j]], rs.uniform(-np.pi, np.pi)
        trajs = np.zeros((13, max(frames - 1, 1), 2), np.float32)
        trajs[0, :, 0] = 0.5
        trajs[1:] = rs.uniform(-0.4, 0.4, (12, trajs.shape[1], 2))
        out.append(dict(gt_boxes=boxes, gt_names=names, gt_fut_trajs=trajs))
    return out


def cpu_reference_run(steps: int, warmup: int, B: int):
    """The reference's CPU path (oracle port of lidargen EfficientUNet + DDIM update, fp32, all host threads).
    /root/reference is not available on the GPU box, so the port (pinned to the reference by tests/golden) is timed."""
    from oracle import unet_torch as O
    import lidarcrafter_b200 as L
    torch.set_grad_enabled(False)
    m = L.EfficientUNet(in_channels=2, resolution=RES, base_channels=64, channel_multiplier=(1, 2, 4, 8),
                        num_residual_blocks=NRES, gn_num_groups=8, gn_eps=1e-6, attn_num_heads=8,
                        coords_encoding="fourier_features", ring=True)
    m.coords = L.get_linear_ray_angles(RES[0], RES[1], 10, -30)
    sd = O.randomize_state_dict(m.state_dict(), seed=0)
    cfg = O.EfficientUNetCfg(resolution=RES, num_residual_blocks=NRES)
    x = torch.randn(B, 2, *RES, generator=torch.Generator().manual_seed(0))
    ts = torch.linspace(1.0, 0.0, 51)
    # thread count: the box is shared and torch's default (all logical cores) oversubscribes badly
    # (measured: 128 threads 38 s vs 16 threads 0.40 s for a batch-2 forward) -> pick the best of a short probe AT THE BATCH
    # SIZE OF THE RUN (a batch-1 probe picked 8 threads on some boxes where 16 were faster at batch 8)
    best = None
    lt_probe = O.log_snr_cosine(ts[:1].repeat(B))
    O.efficient_unet_forward(sd, x[:1], O.log_snr_cosine(ts[:1]), cfg)        # first touch: allocator, weight layout caches
    for n in sorted({8, 16, 32, min(64, os.cpu_count() or 8)}):
        torch.set_num_threads(n)
        t0 = time.perf_counter()
        O.efficient_unet_forward(sd, x, lt_probe, cfg)
        dt = time.perf_counter() - t0
        if best is None or dt < best[0]:
            best = (dt, n)
    torch.set_num_threads(best[1])

    def step(i, x):
        lt, ls = O.log_snr_cosine(ts[i].repeat(B)), O.log_snr_cosine(ts[i + 1].repeat(B))
        return O.ddim_update(x, O.efficient_unet_forward(sd, x, lt, cfg), lt, ls)
    for i in range(warmup):
        x = step(i, x)
    t0 = time.perf_counter()
    for i in range(steps):
        x = step(warmup + i, x)
    dt = time.perf_counter() - t0
    return B * steps / dt, dt, torch.get_num_threads()


def conv_shape(name, a):
    """(H, W, Cin, Cout, taps, bn, rows) of a conv launch of the plan (b200_conv_tc / b200_conv_gn_tc argument lists)"""
    if name == "conv_tc":
        return a[9], a[10], a[11], a[12], a[13], a[15], a[16]
    if name == "conv_tc_splitk":        # + workspace, splits in front of the dimensions
        return a[11], a[12], a[13], a[14], a[15], a[17], a[18]
    if name == "conv_gn_tc":
        return a[21], a[22], a[1] + a[3], a[23], a[24], a[26], a[27]
    return None


def run_temporal(args, dev, rank, world, local):
    """--workload clip | rollout: the composed generation loop (lidarcrafter_b200.temporal.TemporalSampler.generate =
    tools/evaluation/sample_and_save_temporal.py:203-333): first frame from the box-layout model, then the autoregressive
    frames with the device-resident glue in between; one all-gather of the clips at the end."""
    import torch.distributed as dist
    from lidarcrafter_b200.dist import generate_sharded, shard_range
    rollout = args.workload == "rollout"
    frames = args.frames or (20 if rollout else 5)
    total = 8 if rollout else 4 * world                 # configs[4]: batch 8 in total (strong scaling); configs[2]: 4 per GPU
    lo, hi = shard_range(total, rank, world)
    K, W = args.steps, args.warmup
    ts, models = build_temporal(dev, args.precision)
    scenes = synth_scenes(total, frames)
    Bl = hi - lo

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def run(nf, k):          # the public multi-GPU call: rank r generates its clips, one all-gather at the end
        gens = [torch.Generator(device=dev).manual_seed(100 + i) for i in range(total)]
        return generate_sharded(ts, scenes, num_frames=nf, num_steps=k, rng=gens, mode="ddim", temporal_mode="ddim")
    run(min(frames, 2), max(W, 1))                      # warm-up: plans, tile tuning, step graphs of both models
    barrier()
    clk = ClockSampler(local)
    if rank == 0:
        clk.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    clips = run(frames, K)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    clocks = clk.stop() if rank == 0 else None
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms)
    assert clips.shape == (total, frames, 5, *RES) and bool(torch.isfinite(clips).all())
    # e2e: the same public call with the result read back to the host (the reference writes every frame to disk)
    barrier()
    t0 = time.perf_counter()
    host = run(frames, K).cpu()
    torch.cuda.synchronize(dev)
    ms_e = (time.perf_counter() - t0) * 1e3
    if world > 1:
        tms = torch.tensor([ms_e], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms_e = float(tms)
    if rank == 0:
        n_steps = total * frames * K
        plans = [mm.model.get_plan(Bl) for mm in models] if Bl else []
        line = {"metric": METRIC, "value": n_steps / (ms / 1e3), "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
                "ms_per_step": ms / (frames * K), "higher_is_better": True, "scaling": "strong" if rollout else "weak",
                "vs_baseline": None, "dtype": DTYPE_NOTE[args.precision], "data": "synthetic",
                "config": {"workload": ("configs[4]: tri-branch autoregressive rollout, %d frames, batch %d over %d GPU(s)" if rollout else
                                        "configs[2]: layout-conditioned %d-frame clip, batch %d over %d GPU(s)") % (frames, total, world) +
                                       f", LayoutUnetV1 + LayoutTransformerEncoder, {K}-step DDIM per frame, 12 boxes per scene",
                           "frames": frames, "global_batch": total, "batch_per_gpu": Bl, "resolution": list(RES),
                           "l2": "per-step working set of activations >> 126 MB L2"},
                "clocks": clocks, "seconds_per_clip_batch": ms / 1e3,
                "e2e": {"value": n_steps / (ms_e / 1e3), "unit": UNIT, "h2d_bytes_per_step": 0,
                        "d2h_bytes_per_step": int(host.numel() * 4 / (frames * K)),
                        "note": "generate() + device->host copy of the clips, wall clock; conditioning (13 boxes per frame) is uploaded inside the call"},
                "gpu_launches": int(sum(p.plan.n_kernels + 1 for p in plans) / max(len(plans), 1) * frames * K),
                "kernels_per_step": [p.plan.n_kernels + 1 for p in plans]}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--precision", default=DEFAULT_PRECISION, choices=["fp16x3", "fp16f8", "fp16"])
    ap.add_argument("--batch-per-gpu", type=int, default=BATCH_PER_GPU,
                    help="frames per GPU (default 8 = BASELINE configs[1]; other values are side measurements)")
    ap.add_argument("--workload", default="frame", choices=["frame", "clip", "rollout"],
                    help="frame = BASELINE configs[1] (headline); clip = configs[2] (5-frame layout-conditioned clip, batch 4 per "
                         "GPU); rollout = configs[4] (20-frame autoregressive rollout, batch 8 in total, split over the GPUs)")
    ap.add_argument("--frames", type=int, default=0, help="frames per clip (default 5 for clip, 20 for rollout)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profile-ops", action="store_true", help="print the per-kernel time table to stderr")
    ap.add_argument("--profiler-range", action="store_true",
                    help="cudaProfilerStart/Stop around the timed region (ncu --profile-from-start off)")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    B = args.batch_per_gpu
    cfg = {"workload": f"configs[1]: EfficientUNet (nuscenes-unet-uncond) 32x1024, DDIM eta=0, batch {B} per GPU",
           "batch_per_gpu": B, "global_batch": B * world, "resolution": list(RES),
           "l2": "per-step working set ~3 GB of activations per GPU >> 126 MB L2 (inputs larger than L2)"}

    if args.impl == "reference":
        if rank != 0:
            return
        k = max(1, min(args.steps, 2))
        w = 1 if args.warmup > 0 else 0
        v, dt, cores = cpu_reference_run(k, w, B)
        # "steps" / "warmup" are what was TIMED (a bounded sample: the CPU path needs ~2 s per batch-8 step), the requested
        # values ride along
        line = {"impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": k,
                "warmup": w, "steps_requested": args.steps, "warmup_requested": args.warmup, "ms_per_step": 1e3 * dt / k, "higher_is_better": True, "scaling": "weak",
                "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
                "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                 "sample": f"{k} DDIM step(s) at batch {B} after {w} warm-up (oracle port of the reference "
                                           "CPU path; /root/reference does not travel to the GPU box)"},
                "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        print(json.dumps(line))
        return

    import torch.distributed as dist
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.set_grad_enabled(False)
    if args.workload != "frame":
        return run_temporal(args, dev, rank, world, local)
    from lidarcrafter_b200.dist import sample_sharded
    m, ddpm = build_model(dev, args.precision)
    plan = m.get_plan(B)
    K, W = args.steps, args.warmup
    steps = torch.linspace(1.0, 0.0, 51, device=dev)
    x0 = torch.randn(B, 2, *RES, device=dev, generator=torch.Generator(device=dev).manual_seed(rank))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    # the timed call is the PUBLIC sampler: dist.sample_sharded -> GaussianDiffusion.sample(batch_size, num_steps, mode="ddim")
    # (continuous_time.py:236-260: initial noise, per-step tables, per-step noise draw, K denoiser steps) + the path's single
    # collective at the end (all-gather of the final frames, 16 MiB at batch 64)
    gens = torch.Generator(device=dev).manual_seed(100 + rank)     # one generator per rank (a per-sample list draws B times per step)
    sample_sharded(ddpm, B * world, num_steps=max(W, 1), rng=gens, mode="ddim")          # W untimed warm-up steps
    barrier()
    clk = ClockSampler(local)
    if rank == 0:
        clk.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    if args.profiler_range:
        torch.cuda.profiler.start()
    e0.record()
    x_final = sample_sharded(ddpm, B * world, num_steps=K, rng=gens, mode="ddim")         # exactly K timed steps
    e1.record()
    barrier()
    if args.profiler_range:
        torch.cuda.profiler.stop()
    assert x_final.shape[0] == B * world and bool(torch.isfinite(x_final).all())
    ms = e0.elapsed_time(e1)
    clocks = clk.stop() if rank == 0 else None
    if world > 1:
        tms = torch.tensor([ms], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms = float(tms)
    value = B * world * K / (ms / 1e3)

    # ---- e2e: the public API (GaussianDiffusion.p_step) with HOST buffers: every step H2D x_t (pinned) -> p_step -> D2H x_s ----
    xh = torch.empty(B, 2, *RES, pin_memory=True).copy_(x0.cpu())
    yh = torch.empty(B, 2, *RES, pin_memory=True)
    Ke = min(K, 20)
    barrier()
    t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
    xd = torch.empty(B, 2, *RES, device=dev)
    ddpm.p_step(xd.copy_(xh), steps[0].repeat(B), steps[1].repeat(B), mode="ddim")   # untimed warm-up of the public call
    ts = [(steps[(W + i) % 50].repeat(B), steps[(W + i) % 50 + 1].repeat(B)) for i in range(Ke)]   # per-sample step tensors
    barrier()
    t0.record()
    for i in range(Ke):
        xd.copy_(xh, non_blocking=True)                                             # H2D of this step's input
        y = ddpm.p_step(xd, ts[i][0], ts[i][1], mode="ddim")                        # the public call (continuous_time.py:195)
        yh.copy_(y, non_blocking=True)                                              # D2H of this step's result
        torch.cuda.current_stream(dev).synchronize()   # the caller reads the result every step
        xh, yh = yh, xh
    t1.record()
    barrier()
    ms_e = t0.elapsed_time(t1)
    if world > 1:
        tms = torch.tensor([ms_e], device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        ms_e = float(tms)
    e2e_v = B * world * Ke / (ms_e / 1e3)
    nbytes = B * 2 * RES[0] * RES[1] * 4

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- per-kernel timing (CUDA events on the launching stream) for the roofline of the dominant kernel ----
    pk = peaks()
    prof = plan.plan.profile(torch.cuda.current_stream(dev).cuda_stream, reps=3)
    by = {}
    for name, ms_k, fl, by_k in prof:
        d = by.setdefault(name, [0.0, 0.0, 0.0, 0])
        d[0] += ms_k; d[1] += fl; d[2] += by_k; d[3] += 1
    tot_ms = sum(v[0] for v in by.values())
    if args.profile_ops:
        shapes = {}
        for (fn, a), (name, ms_k, fl, by_k) in zip(plan.plan.ops, prof):
            if conv_shape(name, a):
                key = ("conv%s %2dx%-4d Cin%-4d Cout%-4d taps%d bn%d R%d" % ((("+gn" if name == "conv_gn_tc" else "   "),) + conv_shape(name, a)))
                d = shapes.setdefault(key, [0.0, 0.0, 0]); d[0] += ms_k; d[1] += fl; d[2] += 1
            elif name == "gn_act_f16":  # args: x0, C0, x1, C1, st0, st1, g, b, ada, stride, groups, eps, silu, y, parts, B, H, W
                key = "gn_act %2dx%-4d C%-4d norm%d raw%d" % (a[17], a[18], a[1] + a[3], 1 if a[4] else 0, 1 if a[14] else 0)
                d = shapes.setdefault(key, [0.0, 0.0, 0]); d[0] += ms_k; d[1] += by_k / 1e3; d[2] += 1
        # in-kernel wait breakdown of the conv launches inside the real plan (b200_conv_set_debug counters)
        lib = plan.lib
        dbg = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
        waits = {}
        st = torch.cuda.current_stream(dev).cuda_stream
        for (fn, a), (name, ms_k, fl, by_k) in zip(plan.plan.ops, prof):
            if not conv_shape(name, a):
                continue
            dbg.zero_()
            lib.conv_set_debug(dbg.data_ptr())
            fn(*a, st)
            torch.cuda.synchronize(dev)
            lib.conv_set_debug(0)
            d = dbg.view(148, 8).double()
            d = d[d[:, 0] > 0]
            key = ("conv%s %2dx%-4d Cin%-4d Cout%-4d taps%d bn%d R%d" % ((("+gn" if name == "conv_gn_tc" else "   "),) + conv_shape(name, a)))
            w = waits.setdefault(key, [0.0] * 5)
            w[0] += float(d[:, 0].mean()); w[1] += float(d[:, 1].mean()); w[2] += float(d[:, 2].mean())
            w[3] += float(d[:, 3].mean()); w[4] += float(d[:, 5].mean())
        for key, v in sorted(shapes.items(), key=lambda kv: -kv[1][0]):
            if key in waits:
                w = waits[key]
                key = key + "  [mma-thread waits: A %.0f%% B %.0f%% acc %.0f%%; epi idle %.0f%%]" % (
                    100 * w[1] / w[0], 100 * w[2] / w[0], 100 * w[3] / w[0], 100 * w[4] / w[0])
            print(f"    {key}  n={v[2]:2d} {v[0]:7.3f} ms  {v[1] / max(v[0], 1e-9) / 1e9:7.1f} TFLOP/s | TB/s (algorithmic)", file=sys.stderr)
        for name, v in sorted(by.items(), key=lambda kv: -kv[1][0]):
            print(f"  {name:20s} n={v[3]:4d} {v[0]:8.3f} ms  {100 * v[0] / tot_ms:5.1f}%  "
                  f"{v[1] / max(v[0], 1e-9) / 1e9:8.1f} TFLOP/s  {v[2] / max(v[0], 1e-9) / 1e6:8.1f} GB/s", file=sys.stderr)
    c = [sum(by.get(n, [0, 0, 0, 0])[i] for n in ("conv_tc", "conv_tc_splitk", "conv_gn_tc")) for i in range(4)]
    # tensor-pipe time units per algorithmic product (1 unit = one fp16 MMA; the e4m3 K=32 correction MMA of fp16f8 = 1)
    mma_per_product = {"fp16x3": 3, "fp16f8": 2, "fp16": 1}[args.precision]
    # dominant kernel = the conv shape with the largest share of the step (per-launch numbers)
    dom = {}
    for (fn, a), (name, ms_k, fl, by_k) in zip(plan.plan.ops, prof):
        if conv_shape(name, a):
            key = conv_shape(name, a)[:5]
            d = dom.setdefault(key, [0.0, 0.0, 0.0, 0]); d[0] += ms_k; d[1] += fl; d[2] += by_k; d[3] += 1
    dk, dv = max(dom.items(), key=lambda kv: kv[1][0])
    d_ms, d_fl, d_by = dv[0] / dv[3], dv[1] / dv[3], dv[2] / dv[3]
    t_tensor = d_fl / (pk["tf"] * 1e12) * 1e3          # ms at the measured bf16 peak (1 MMA / product)
    t_hbm = d_by / (pk["hbm_gbs"] * 1e9) * 1e3         # ms at the measured HBM peak
    bound = "hbm" if t_hbm >= t_tensor else "tensor"
    ach = (d_by / (d_ms / 1e3) / 1e9) if bound == "hbm" else (d_fl / (d_ms / 1e3) / 1e12)
    peak = pk["hbm_gbs"] if bound == "hbm" else pk["tf"]
    # DRAM traffic of this shape from the committed ncu --set full capture (profiles/), per launch; None if unknown
    # (profiles/r01s2_ncu_summary.md: 101.8 MB without / 182.2 MB with the residual read, 6 launches each per step; the writes
    # of a launch that are still in L2 when it ends are not in dram__bytes_write)
    dom_names = sorted({("conv_col_kernel (fused GroupNorm+SiLU front end, column walk)" if conv_shape(name, a)[6] == 0 else
                         "conv_tc_kernel (fused front end)" if name == "conv_gn_tc" else "conv_tc_kernel")
                        for (fn, a), (name, _, _, _) in zip(plan.plan.ops, prof) if conv_shape(name, a) and conv_shape(name, a)[:5] == dk})
    ncu_traffic = {(32, 1024, 64, 64, 9): NCU_TRAFFIC_32x1024_C64}.get(dk)
    if ncu_traffic is not None and any("conv_col" in n for n in dom_names):
        # the column-walk kernel: mean DRAM bytes per launch from ITS capture (tools/ncu_traffic.py on the .ncu-rep of
        # `ncu --set full -k regex:conv_col` over this bench command; profiles/r02b_col_traffic.json)
        for path in ("gpurun_out/r02b_col_traffic.json", "profiles/r02b_col_traffic.json"):
            full = os.path.join(os.path.dirname(os.path.abspath(__file__)), path)
            if os.path.exists(full):
                ncu_traffic = json.load(open(full))["traffic_bytes"]
                break
        else:
            ncu_traffic = None
    roof = {"bound": bound, "kernel": "%s %dx%d Cin%d Cout%d taps%d (x%d launches/step, %.0f%% of the step)" % (
                (" | ".join(dom_names),) + dk + (dv[3], 100 * dv[0] / tot_ms)),
            "achieved": ach, "peak": peak, "unit": "GB/s" if bound == "hbm" else "TFLOP/s", "frac": ach / peak,
            "traffic": ncu_traffic, "peak_source": pk["src"],
            "per_launch": {"ms": d_ms, "algorithmic_gflop": d_fl / 1e9, "algorithmic_mb": d_by / 1e6,
                           "ms_at_tensor_peak": t_tensor, "ms_at_hbm_peak": t_hbm,
                           "tflops_algorithmic": d_fl / (d_ms / 1e3) / 1e12,
                           "tensor_pipe_tflops": mma_per_product * d_fl / (d_ms / 1e3) / 1e12},
            "all_conv_launches": {"n": c[3], "ms": c[0], "share_of_step": c[0] / tot_ms,
                                  "tflops_algorithmic": c[1] / (c[0] / 1e3) / 1e12,
                                  "frac_of_bf16_peak_algorithmic": c[1] / (c[0] / 1e3) / 1e12 / pk["tf"],
                                  "tensor_work_multiplier": mma_per_product},
            "note": "bound = max(algorithmic FLOPs / measured bf16 peak, algorithmic bytes / measured HBM peak) per launch "
                    "(SURVEY 8d); tensor_pipe_tflops counts the MMAs really issued per algorithmic product "
                    "(fp16x3: 3 fp16; fp16f8: 1 fp16 + 1 e4m3 at K=32; fp16: 1)"}
    # ---- the north_star's unit: the ResBlock (norm1 -> SiLU -> conv1 -> AdaGN -> SiLU -> conv2 (+1x1 skip) + residual), ALL of
    # its launches; algorithmic bytes / FLOPs per SURVEY 8d: 4 B (2 Cin + 3 Cout) per pixel + weights, 2 (9 Cin Cout + 9 Cout^2
    # [+ Cin Cout]) per pixel ----
    blocks = {}
    for tag, (name, ms_k, fl, by_k) in zip(plan.plan.tags, prof):
        if tag.startswith("resblock"):
            shape = tag.rsplit(" #", 1)[0]
            d = blocks.setdefault(shape, {"ms": 0.0, "launches": 0, "ids": set()})
            d["ms"] += ms_k; d["launches"] += 1; d["ids"].add(tag)
    if blocks:
        shape, d = max(blocks.items(), key=lambda kv: kv[1]["ms"])
        hw, chans = shape.split()[1], shape.split()[2]
        h_, w_ = (int(v) for v in hw.split("x"))
        ci, co = (int(v) for v in chans.split("->"))
        skip = shape.endswith("skip")
        nb = len(d["ids"])
        by_blk = 4.0 * B * h_ * w_ * (2 * ci + 3 * co) + 4.0 * 9 * (ci + co) * co
        fl_blk = 2.0 * B * h_ * w_ * (9 * ci * co + 9 * co * co + (ci * co if skip else 0))
        ms_blk = d["ms"] / nb
        t_h, t_t = by_blk / (pk["hbm_gbs"] * 1e9) * 1e3, fl_blk / (pk["tf"] * 1e12) * 1e3
        roof["resblock"] = {"shape": shape, "blocks_per_step": nb, "launches_per_block": d["launches"] / nb, "ms_per_block": ms_blk,
                            "share_of_step": d["ms"] / tot_ms, "algorithmic_mb": by_blk / 1e6, "algorithmic_gflop": fl_blk / 1e9,
                            "bound": "hbm" if t_h >= t_t else "tensor", "ms_at_peak": max(t_h, t_t),
                            "achieved_gbs": by_blk / (ms_blk / 1e3) / 1e9, "achieved_tflops": fl_blk / (ms_blk / 1e3) / 1e12,
                            "frac": max(t_h, t_t) / ms_blk}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": DTYPE_NOTE[args.precision],
            "data": "synthetic", "config": cfg, "clocks": clocks,
            "e2e": {"value": e2e_v, "unit": UNIT, "h2d_bytes_per_step": nbytes, "d2h_bytes_per_step": nbytes,
                    "steps": Ke},
            "gpu_launches": K * (plan.plan.n_kernels + 2), "kernels_per_step": plan.plan.n_kernels + 1,
            "timed_call": "dist.sample_sharded -> ContinuousTimeGaussianDiffusion.sample(batch_size, num_steps=K, mode='ddim')",
            "roofline": roof,
            "algorithmic_gflop_per_sample_step": plan.plan.flops / B / 1e9}
    if not args.no_cpu_baseline and world == 1:      # reported on rank 0 at N = 1 only
        v, dt, cores = cpu_reference_run(1, 1, B)
        line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": f"1 DDIM step at batch {B} after 1 warm-up step ({dt:.1f} s)"}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
