"""fp32 CPU restatement of the reference range-image denoiser (EfficientUNet) -- TEST ORACLE.

Functional PyTorch (no nn.Module): every function takes the reference ``state_dict`` (same key
names as ``lidargen/models/unets/efficient_unet.py``) and plain tensors in the reference's NCHW
layout.  This is the "torch fp32 reference" for the floating-point kernels; it is pinned against
the UNMODIFIED reference through ``tests/golden`` (see ``tests/test_oracle_golden.py``).

Only tests / smoke / bench's CPU-baseline leg may import this file.
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F


@dataclass
class EfficientUNetCfg:
    """ctor kwargs of the reference EfficientUNet (efficient_unet.py:199-213)."""
    in_channels: int = 2
    resolution: tuple = (32, 1024)
    out_channels: int | None = None
    base_channels: int = 64
    temb_channels: int | None = None
    channel_multiplier: tuple = (1, 2, 4, 8)
    num_residual_blocks: tuple = (3, 3, 3, 3)
    gn_num_groups: int = 8
    gn_eps: float = 1e-6
    attn_num_heads: int = 8
    coords_encoding: str | None = "fourier_features"
    ring: bool = True

    def __post_init__(self):
        if self.out_channels is None:
            self.out_channels = self.in_channels
        if self.temb_channels is None:
            self.temb_channels = self.base_channels * 4
        if isinstance(self.channel_multiplier, int):
            self.channel_multiplier = (self.channel_multiplier,) * 4
        if isinstance(self.num_residual_blocks, int):
            self.num_residual_blocks = (self.num_residual_blocks,) * 4


# ---------------------------------------------------------------------------------------------
# building blocks (reference: lidargen/models/unets/ops.py)
# ---------------------------------------------------------------------------------------------
def ring_pad(x: torch.Tensor, pad: int, ring: bool) -> torch.Tensor:
    """ops.py:32-43 -- circular in W when ``ring`` (else zeros), zeros in H."""
    if pad == 0:
        return x
    x = F.pad(x, (pad, pad, 0, 0), mode="circular" if ring else "constant")
    return F.pad(x, (0, 0, pad, pad), mode="constant")


def conv2d_ring(x, w, b, ring: bool):
    """ops.py:149-173 -- k x k conv, padding k//2 done by ring_pad (1x1 convs are unpadded)."""
    k = w.shape[-1]
    return F.conv2d(ring_pad(x, k // 2, ring), w, b)


def sinusoidal_embedding(t: torch.Tensor, channels: int, max_period: int = 10_000) -> torch.Tensor:
    """ops.py:14-26."""
    half = channels // 2
    f = torch.exp(-math.log(max_period) / (half - 1) * torch.arange(half, dtype=torch.float32))
    a = t[:, None].float() * f[None, :]
    return torch.cat([a.sin(), a.cos()], dim=-1)


def fir_down2(x: torch.Tensor, ring: bool) -> torch.Tensor:
    """ops.py:52-146 with down=2, window [1,3,3,1]/8 (closed form):
    out[i] = (x[2i-1] + 3 x[2i] + 3 x[2i+1] + x[2i+2]) / 8, W circular / H zero, W first then H."""
    def axis(v, dim, circ):
        n = v.shape[dim]
        if circ:
            vm1 = torch.roll(v, 1, dim)
            vp1 = torch.roll(v, -1, dim)
            vp2 = torch.roll(v, -2, dim)
        else:
            z = torch.zeros_like(v.narrow(dim, 0, 1))
            vm1 = torch.cat([z, v.narrow(dim, 0, n - 1)], dim)
            vp1 = torch.cat([v.narrow(dim, 1, n - 1), z], dim)
            vp2 = torch.cat([v.narrow(dim, 2, n - 2), z, z], dim)
        full = (vm1 + 3 * v + 3 * vp1 + vp2) / 8
        idx = torch.arange(0, n, 2)
        return full.index_select(dim, idx)
    x = axis(x, 3, ring)
    x = axis(x, 2, False)
    return x


def fir_up2(x: torch.Tensor, ring: bool) -> torch.Tensor:
    """ops.py:52-146 with up=2 (gain 2 per axis):
    out[2i] = (x[i-1] + 3 x[i]) / 4, out[2i+1] = (3 x[i] + x[i+1]) / 4."""
    def axis(v, dim, circ):
        n = v.shape[dim]
        if circ:
            vm1 = torch.roll(v, 1, dim)
            vp1 = torch.roll(v, -1, dim)
        else:
            z = torch.zeros_like(v.narrow(dim, 0, 1))
            vm1 = torch.cat([z, v.narrow(dim, 0, n - 1)], dim)
            vp1 = torch.cat([v.narrow(dim, 1, n - 1), z], dim)
        even = (vm1 + 3 * v) / 4
        odd = (3 * v + vp1) / 4
        out = torch.stack([even, odd], dim=dim + 1)
        shape = list(v.shape)
        shape[dim] = 2 * n
        return out.reshape(shape)
    x = axis(x, 3, ring)
    x = axis(x, 2, False)
    return x


def fourier_features(coords: torch.Tensor, resolution) -> torch.Tensor:
    """encoding.py:120-149 -- coords [1,2,H,W] (elev, azim) -> [1, 2*(L_h+L_w), H, W]."""
    L_h = int(math.ceil(math.log2(resolution[0])))
    L_w = int(math.ceil(math.log2(resolution[1])))
    fh = torch.cat([torch.arange(L_h).float().exp2(), torch.zeros(L_w)])
    fw = torch.cat([torch.zeros(L_h), torch.arange(L_w).float().exp2()])
    ang = coords[:, 0:1] * fh[None, :, None, None] + coords[:, 1:2] * fw[None, :, None, None]
    return torch.cat([ang.sin(), ang.cos()], dim=1)


def linear_ray_angles(H: int, W: int, fov_up: float, fov_down: float) -> torch.Tensor:
    """lidargen/utils/lidar.py:22-32 -- [1,2,H,W] elevation / azimuth in radians."""
    elev = (1 - torch.arange(H) / H) * (fov_up - fov_down) + fov_down
    azim = (1 - torch.arange(W) / W) * 360.0 - 180.0
    e, a = torch.meshgrid([elev, azim], indexing="ij")
    return torch.stack([e, a])[None].deg2rad()


# ---------------------------------------------------------------------------------------------
# EfficientUNet (reference: lidargen/models/unets/efficient_unet.py)
# ---------------------------------------------------------------------------------------------
def _resblock(sd, p, x, temb, cfg: EfficientUNetCfg):
    """efficient_unet.py:61-115 (ResidualBlock) + ops.py:176-200 (AdaGN)."""
    G, eps = cfg.gn_num_groups, cfg.gn_eps
    h = F.group_norm(x, G, sd[p + "norm1.weight"], sd[p + "norm1.bias"], eps)
    h = F.silu(h)
    h = conv2d_ring(h, sd[p + "conv1.weight"], sd[p + "conv1.bias"], cfg.ring)
    h = F.group_norm(h, G, None, None, eps)
    ss = F.linear(F.silu(temb), sd[p + "norm2.proj.1.weight"], sd[p + "norm2.proj.1.bias"])
    scale, shift = ss.chunk(2, dim=1)
    h = h * (1 + scale[:, :, None, None]) + shift[:, :, None, None]
    h = F.silu(h)
    h = conv2d_ring(h, sd[p + "conv2.weight"], sd[p + "conv2.bias"], cfg.ring)
    if (p + "skip.weight") in sd:
        x = F.conv2d(x, sd[p + "skip.weight"], sd[p + "skip.bias"])
    return (x + h) * sd[p + "scale"]


def _self_attention(sd, p, x, cfg: EfficientUNetCfg):
    """efficient_unet.py:28-58 (GroupNorm -> nn.MultiheadAttention over H*W tokens)."""
    B, C, H, W = x.shape
    nh = cfg.attn_num_heads
    d = C // nh
    h = F.group_norm(x, cfg.gn_num_groups, sd[p + "norm.weight"], sd[p + "norm.bias"], cfg.gn_eps)
    tok = h.flatten(2).transpose(1, 2)  # B, T, C
    qkv = F.linear(tok, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"])
    q, k, v = qkv.chunk(3, dim=-1)
    def heads(t):
        return t.reshape(B, -1, nh, d).transpose(1, 2)  # B, nh, T, d
    q, k, v = heads(q), heads(k), heads(v)
    att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(d), dim=-1)
    o = (att @ v).transpose(1, 2).reshape(B, -1, C)
    o = F.linear(o, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])
    o = o.transpose(1, 2).reshape(B, C, H, W)
    return (x + o) * sd[p + "scale"]


def _block(sd, p, h, temb, cfg, n_res, down=False, up=False, attn=False):
    """efficient_unet.py:118-190 (Block)."""
    if down:
        h = conv2d_ring(h, sd[p + "downsample.0.weight"], sd[p + "downsample.0.bias"], cfg.ring)
        h = fir_down2(h, cfg.ring)
    for i in range(n_res):
        h = _resblock(sd, f"{p}residual_blocks.{i}.", h, temb, cfg)
    if attn:
        h = _self_attention(sd, p + "self_attn_block.", h, cfg)
    if up:
        h = fir_up2(h, cfg.ring)
        h = conv2d_ring(h, sd[p + "upsample.1.weight"], sd[p + "upsample.1.bias"], cfg.ring)
    return h


def time_embedding(sd, t, cfg: EfficientUNetCfg):
    """efficient_unet.py:237-242."""
    e = sinusoidal_embedding(t, cfg.base_channels)
    e = F.linear(e, sd["time_embedding.1.weight"], sd["time_embedding.1.bias"])
    e = F.silu(e)
    return F.linear(e, sd["time_embedding.3.weight"], sd["time_embedding.3.bias"])


def efficient_unet_forward(sd: dict, images: torch.Tensor, timesteps: torch.Tensor,
                           cfg: EfficientUNetCfg) -> torch.Tensor:
    """efficient_unet.py:274-300."""
    B = images.shape[0]
    if timesteps.dim() == 0:
        timesteps = timesteps[None].repeat_interleave(B, dim=0)
    temb = time_embedding(sd, timesteps.float(), cfg)
    h = images
    if cfg.coords_encoding == "fourier_features":
        cenc = fourier_features(sd["coords"], cfg.resolution).repeat_interleave(B, dim=0)
        h = torch.cat([h, cenc], dim=1)
    elif cfg.coords_encoding is not None:
        raise NotImplementedError(cfg.coords_encoding)
    N = cfg.num_residual_blocks
    h = conv2d_ring(h, sd["in_conv.weight"], sd["in_conv.bias"], cfg.ring)
    h1 = _block(sd, "d_block1.", h, temb, cfg, N[0])
    h2 = _block(sd, "d_block2.", h1, temb, cfg, N[1], down=True)
    h3 = _block(sd, "d_block3.", h2, temb, cfg, N[2], down=True)
    h4 = _block(sd, "d_block4.", h3, temb, cfg, N[3], down=True, attn=True)
    h = _block(sd, "u_block4.", h4, temb, cfg, N[3], up=True, attn=True)
    h = _block(sd, "u_block3.", torch.cat([h, h3], 1), temb, cfg, N[2], up=True)
    h = _block(sd, "u_block2.", torch.cat([h, h2], 1), temb, cfg, N[1], up=True)
    h = _block(sd, "u_block1.", torch.cat([h, h1], 1), temb, cfg, N[0])
    return conv2d_ring(h, sd["out_conv.weight"], sd["out_conv.bias"], cfg.ring)


# ---------------------------------------------------------------------------------------------
# continuous-time diffusion (reference: lidargen/models/diffusion/continuous_time.py)
# ---------------------------------------------------------------------------------------------
def log_snr_cosine(t: torch.Tensor, logsnr_min: float = -15.0, logsnr_max: float = 15.0):
    """continuous_time.py:22-29."""
    t_min = math.atan(math.exp(-0.5 * logsnr_max))
    t_max = math.atan(math.exp(-0.5 * logsnr_min))
    return -2 * torch.log(torch.tan(t_min + t * (t_max - t_min)).clamp(min=1e-20))


def alpha_sigma(log_snr: torch.Tensor):
    """continuous_time.py:61-63."""
    return log_snr.sigmoid().sqrt(), (-log_snr).sigmoid().sqrt()


def ddim_update(x_t, pred, log_snr_t, log_snr_s, noise=None, eta: float = 0.0,
                objective: str = "eps", clip: float | None = 1.0):
    """continuous_time.py:205-231 (mode='ddim').  log_snr_* are [B]."""
    lt = log_snr_t[:, None, None, None]
    ls = log_snr_s[:, None, None, None]
    a_t, s_t = alpha_sigma(lt)
    a_s, s_s = alpha_sigma(ls)
    if objective == "eps":
        x0 = (x_t - s_t * pred) / a_t
    elif objective == "v":
        x0 = a_t * x_t - s_t * pred
    else:
        x0 = pred
    if clip is not None:
        x0 = x0.clamp(-clip, clip)
    c1 = eta * s_s / s_t * (1 - a_t ** 2 / a_s ** 2).sqrt()
    c2 = (1 - a_s ** 2 - c1 ** 2).sqrt()
    eps = (x_t - a_t * x0) / s_t
    out = a_s * x0 + c2 * eps
    if noise is not None:
        out = out + c1 * noise
    return out


def ddpm_update(x_t, pred, log_snr_t, log_snr_s, noise, objective: str = "eps",
                clip: float | None = 1.0):
    """continuous_time.py:205-225 (mode='ddpm')."""
    lt = log_snr_t[:, None, None, None]
    ls = log_snr_s[:, None, None, None]
    a_t, s_t = alpha_sigma(lt)
    a_s, s_s = alpha_sigma(ls)
    if objective == "eps":
        x0 = (x_t - s_t * pred) / a_t
    elif objective == "v":
        x0 = a_t * x_t - s_t * pred
    else:
        x0 = pred
    if clip is not None:
        x0 = x0.clamp(-clip, clip)
    c = -torch.expm1(lt - ls)
    mean = a_s * (x_t * (1 - c) / a_t + c * x0)
    return mean + s_s * c.sqrt() * noise


def sample_uncond(model_fn, x_T: torch.Tensor, num_steps: int, mode: str = "ddim", eta: float = 0.0,
                  noises=None, objective: str = "eps", return_all: bool = False):
    """continuous_time.py:237-260.  ``model_fn(x, log_snr[B])``; ``noises`` = per-step noise list
    (the reference draws one randn per step even for eta == 0)."""
    B = x_T.shape[0]
    steps = torch.linspace(1.0, 0.0, num_steps + 1)
    x = x_T
    out = [x]
    for i in range(num_steps):
        lt = log_snr_cosine(steps[i].repeat(B))
        ls = log_snr_cosine(steps[i + 1].repeat(B))
        pred = model_fn(x, lt)
        nz = None if noises is None else noises[i]
        if mode == "ddim":
            x = ddim_update(x, pred, lt, ls, nz, eta, objective)
        else:
            x = ddpm_update(x, pred, lt, ls, nz, objective)
        out.append(x)
    return torch.stack(out) if return_all else x


# ---------------------------------------------------------------------------------------------
# deterministic weights shared by the reference, the oracle and the CUDA path
# ---------------------------------------------------------------------------------------------
def randomize_state_dict(sd: dict, seed: int = 0) -> dict:
    """Re-randomise every floating parameter (the reference zero-inits conv2/out_conv/out_proj,
    SURVEY section 4 trap 1).  One generator per key (seeded from the key name) so the values do not
    depend on construction order.  GroupNorm gamma ~ N(1, 0.1), beta ~ N(0, 0.1); biases ~ N(0, 0.02);
    weights ~ N(0, 1/fan_in) so every branch stays O(1).  Buffers (coords, scale, kernel, freqs,
    phase) are left untouched."""
    import zlib
    out = {}
    for k, v in sd.items():
        leaf = k.split(".")[-1]
        is_buf = leaf in ("coords", "scale", "kernel", "freqs", "phase", "_dummy") or not v.is_floating_point()
        if is_buf:
            out[k] = v.clone()
            continue
        g = torch.Generator().manual_seed((zlib.crc32(k.encode()) + 7919 * seed) % (2 ** 31))
        if leaf == "weight" and v.dim() == 1:      # GroupNorm / LayerNorm gamma (any 1-D ``weight``)
            out[k] = 1 + 0.1 * torch.randn(v.shape, generator=g)
        elif "norm" in k and leaf == "bias" and "proj" not in k:
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif leaf in ("bias", "in_proj_bias") or v.dim() == 1:
            out[k] = 0.02 * torch.randn(v.shape, generator=g)
        else:
            fan_in = v[0].numel()
            out[k] = torch.randn(v.shape, generator=g) / math.sqrt(fan_in)
    return out


# =============================================================================================
# LayoutUnetV1 + LayoutTransformerEncoder (layout-conditioned denoiser, configs 3 and 5)
#   reference: lidargen/models/unets/layout_unet_v1.py, layout_encoder.py, nn.py
# =============================================================================================
@dataclass
class LayoutUnetCfg:
    """ctor kwargs of LayoutUnetV1 that matter for inference (option_nusc_box_layout_v3.py:10-33)."""
    in_channels: int = 12
    resolution: tuple = (32, 1024)
    model_channels: int = 64
    out_channels: int = 2
    num_res_blocks: int = 2
    attention_ds: tuple = (4, 8)
    encoder_channels: int = 64
    channel_mult: tuple = (1, 2, 4, 8)
    num_head_channels: int = 32
    image_size: int = 32
    ring: bool = True
    gn_groups: int = 32
    gn_eps: float = 1e-5


def _gn32(x, w, b, cfg):
    """nn.py:17-19,104-111 -- GroupNorm(32, C) computed in fp32 (works for [B,C,L] and [B,C,H,W])."""
    return F.group_norm(x.float(), cfg.gn_groups, w, b, cfg.gn_eps)


def _layout_resblock(sd, p, x, emb, cfg, up=False, down=False):
    """layout_unet_v1.py:143-249 (ResBlock, use_scale_shift_norm=True)."""
    h = F.silu(_gn32(x, sd[p + "in_layers.0.weight"], sd[p + "in_layers.0.bias"], cfg))
    if up:
        h, x = fir_up2(h, cfg.ring), fir_up2(x, cfg.ring)
    elif down:
        h, x = fir_down2(h, cfg.ring), fir_down2(x, cfg.ring)
    h = conv2d_ring(h, sd[p + "in_layers.2.weight"], sd[p + "in_layers.2.bias"], cfg.ring)
    e = F.linear(F.silu(emb), sd[p + "emb_layers.1.weight"], sd[p + "emb_layers.1.bias"])
    scale, shift = e.chunk(2, dim=1)
    h = _gn32(h, sd[p + "out_layers.0.weight"], sd[p + "out_layers.0.bias"], cfg)
    h = h * (1 + scale[:, :, None, None]) + shift[:, :, None, None]
    h = conv2d_ring(F.silu(h), sd[p + "out_layers.3.weight"], sd[p + "out_layers.3.bias"], cfg.ring)
    if (p + "skip_connection.weight") in sd:
        x = F.conv2d(x, sd[p + "skip_connection.weight"], sd[p + "skip_connection.bias"])
    return x + h


def _object_aware_attention(sd, p, x, cond, cfg: LayoutUnetCfg):
    """layout_unet_v1.py:347-532 (norm_first=False, channels_scale_for_positional_embedding=1, no padding mask)."""
    B, C, H, W = x.shape
    nh = C // cfg.num_head_channels
    T = H * W
    xf = x.reshape(B, C, T)
    qkv = F.conv1d(_gn32(xf, sd[p + "norm_for_qkv.weight"], sd[p + "norm_for_qkv.bias"], cfg),
                   sd[p + "qkv_projector.weight"], sd[p + "qkv_projector.bias"])
    res = cfg.image_size // (cfg.resolution[0] // H)
    pw, pb = sd[p + "layout_position_embedding_projector.weight"], sd[p + "layout_position_embedding_projector.bias"]
    pos_p = _gn32(F.conv1d(cond[f"image_patch_bbox_embedding_for_resolution{res}"], pw, pb),
                  sd[p + "norm_for_image_patch_positional_embedding.weight"],
                  sd[p + "norm_for_image_patch_positional_embedding.bias"], cfg)
    pos_l = _gn32(F.conv1d(cond["obj_bbox_embedding"], pw, pb), sd[p + "norm_for_layout_positional_embedding.weight"],
                  sd[p + "norm_for_layout_positional_embedding.bias"], cfg)
    content = (cond["xf_out"] + _gn32(cond["obj_class_embedding"], sd[p + "norm_for_obj_class_embedding.weight"],
                                      sd[p + "norm_for_obj_class_embedding.bias"], cfg)) / 2
    kl, vl = F.conv1d(content, sd[p + "layout_content_embedding_projector.weight"],
                      sd[p + "layout_content_embedding_projector.bias"]).split(C, dim=1)
    L2 = kl.shape[-1]
    d = C // nh

    def hd(t, n):
        return t.reshape(B * nh, d, n)
    q, k, v = [hd(t, T) for t in qkv.split(C, dim=1)]
    pp, pl = hd(pos_p, T), hd(pos_l, L2)
    qm = torch.cat([q, pp], 1)
    km = torch.cat([torch.cat([k, pp], 1), torch.cat([hd(kl, L2), pl], 1)], 2)
    vm = torch.cat([v, hd(vl, L2)], 2)
    scale = 1 / math.sqrt(math.sqrt(2 * d))
    w = torch.softmax(torch.einsum("bct,bcs->bts", qm * scale, km * scale).float(), dim=-1)
    a = torch.einsum("bts,bcs->bct", w, vm).reshape(B, C, T)
    h = F.conv1d(a, sd[p + "proj_out.weight"], sd[p + "proj_out.bias"])
    return (xf + h).reshape(B, C, H, W)


def layout_block_list(cfg: LayoutUnetCfg):
    """Mirror of the constructor's block schedule (layout_unet_v1.py:690-852): returns
    (input_blocks, output_blocks) as lists of lists of (kind, cin, cout) with kind in
    {'conv','res','res_down','res_up','attn'}; also the running skip-channel stack."""
    mc = cfg.model_channels
    ch = int(cfg.channel_mult[0] * mc)
    inputs = [[("conv", None, ch)]]
    chans = [ch]
    ds = 1
    for level, mult in enumerate(cfg.channel_mult):
        for _ in range(cfg.num_res_blocks):
            layers = [("res", ch, int(mult * mc))]
            ch = int(mult * mc)
            if ds in cfg.attention_ds:
                layers.append(("attn", ch, ch))
            inputs.append(layers)
            chans.append(ch)
        if level != len(cfg.channel_mult) - 1:
            inputs.append([("res_down", ch, ch)])
            chans.append(ch)
            ds *= 2
    middle = [("res", ch, ch), ("attn", ch, ch), ("res", ch, ch)]
    outputs = []
    for level, mult in list(enumerate(cfg.channel_mult))[::-1]:
        for i in range(cfg.num_res_blocks + 1):
            ich = chans.pop()
            layers = [("res", ch + ich, int(mc * mult))]
            ch = int(mc * mult)
            if ds in cfg.attention_ds:
                layers.append(("attn", ch, ch))
            if level and i == cfg.num_res_blocks:
                layers.append(("res_up", ch, ch))
                ds //= 2
            outputs.append(layers)
    return inputs, middle, outputs


def layout_unet_forward(sd: dict, x: torch.Tensor, log_snr: torch.Tensor, cond: dict, cfg: LayoutUnetCfg):
    """layout_unet_v1.py:866-902.  ``cond`` = output dict of layout_encoder_forward."""
    B = x.shape[0]
    e = sinusoidal_embedding(log_snr.float(), cfg.model_channels)
    e = F.silu(F.linear(e, sd["time_embed.1.weight"], sd["time_embed.1.bias"]))
    emb = F.linear(e, sd["time_embed.3.weight"], sd["time_embed.3.bias"]) + cond["xf_proj"]
    h = x
    if "concat_cond" in cond:
        h = torch.cat([h, cond["concat_cond"]], dim=1)
    h = torch.cat([h, fourier_features(sd["coords"], cfg.resolution).repeat_interleave(B, dim=0)], dim=1)
    inputs, middle, outputs = layout_block_list(cfg)

    def run(layers, prefix, h):
        for j, (kind, cin, cout) in enumerate(layers):
            p = f"{prefix}{j}."
            if kind == "conv":
                h = conv2d_ring(h, sd[p + "weight"], sd[p + "bias"], cfg.ring)
            elif kind == "attn":
                h = _object_aware_attention(sd, p, h, cond, cfg)
            else:
                h = _layout_resblock(sd, p, h, emb, cfg, up=kind == "res_up", down=kind == "res_down")
        return h
    hs = []
    for i, layers in enumerate(inputs):
        h = run(layers, f"input_blocks.{i}.", h)
        hs.append(h)
    h = run(middle, "middle_block.", h)
    for i, layers in enumerate(outputs):
        h = run(layers, f"output_blocks.{i}.", torch.cat([h, hs.pop()], dim=1))
    h = F.silu(_gn32(h, sd["out.0.weight"], sd["out.0.bias"], cfg))
    return conv2d_ring(h, sd["out.2.weight"], sd["out.2.bias"], cfg.ring)


def layout_encoder_forward(sd: dict, batch: dict, feature_map_size=(32, 1024), resolutions=(4, 8), heads: int = 4,
                           layers: int = 6):
    """layout_encoder.py:237-303 (used_condition_types = obj_class, obj_bbox, is_valid_obj; no positional
    embedding; final LayerNorm; no key padding mask)."""
    obj_bbox = batch["scaled_gt_boxes"][..., :8].float()
    obj_bbox_2d = batch["gt_boxes_2d"].float()
    obj_class = batch["scaled_gt_boxes"][..., -1].long()
    out = {}
    cls_e = F.embedding(obj_class, sd["obj_class_embedding.weight"])
    box_e = F.linear(obj_bbox, sd["obj_bbox_embedding.weight"], sd["obj_bbox_embedding.bias"])
    box2_e = F.linear(obj_bbox_2d, sd["obj_bbox_2d_embedding.weight"], sd["obj_bbox_2d_embedding.bias"])
    x = cls_e + box_e + box2_e
    out["obj_class_embedding"] = cls_e.permute(0, 2, 1)
    out["obj_bbox_embedding"] = box2_e.permute(0, 2, 1)
    Bn = x.shape[0]
    for r in resolutions:
        Hr, Wr = int(feature_map_size[0] / r), int(feature_map_size[1] / r)
        ii, ij = 1.0 / (feature_map_size[0] / r), 1.0 / (feature_map_size[1] / r)
        tab = torch.tensor([(ij * j, ii * i, ij * (j + 1), ii * (i + 1)) for i in range(Hr) for j in range(Wr)],
                           dtype=torch.float32)
        emb = F.linear(tab, sd["obj_bbox_2d_embedding.weight"], sd["obj_bbox_2d_embedding.bias"])
        out[f"image_patch_bbox_embedding_for_resolution{Hr}"] = emb[None].repeat_interleave(Bn, 0).permute(0, 2, 1)
    out["key_padding_mask"] = (1 - batch["is_valid_obj"]).bool()
    width = x.shape[-1]
    ach = width // heads
    sc = 1 / math.sqrt(math.sqrt(ach))
    for l in range(layers):
        p = f"transform.resblocks.{l}."
        y = F.layer_norm(x, (width,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"])
        qkv = F.linear(y, sd[p + "attn.c_qkv.weight"], sd[p + "attn.c_qkv.bias"])
        bs, n, _ = qkv.shape
        q, k, v = qkv.view(bs, n, heads, -1).split(ach, dim=-1)
        w = torch.softmax(torch.einsum("bthc,bshc->bhts", q * sc, k * sc).float(), dim=-1)
        a = torch.einsum("bhts,bshc->bthc", w, v).reshape(bs, n, -1)
        x = x + F.linear(a, sd[p + "attn.c_proj.weight"], sd[p + "attn.c_proj.bias"])
        y = F.layer_norm(x, (width,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"])
        y = F.linear(F.gelu(F.linear(y, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"])),
                     sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
        x = x + y
    x = F.layer_norm(x, (width,), sd["final_ln.weight"], sd["final_ln.bias"])
    out["xf_proj"] = F.linear(x[:, 0], sd["transformer_proj.weight"], sd["transformer_proj.bias"])
    out["xf_out"] = x.permute(0, 2, 1)
    if "concat_cond" in batch:
        if "autoregressive_cond" in batch:
            out["concat_cond"] = torch.cat([batch["concat_cond"], batch["autoregressive_cond"]], dim=1)
        else:
            out["concat_cond"] = batch["concat_cond"]
    return out


def synth_layout_batch(B: int, H: int = 32, W: int = 1024, seed: int = 0, n_obj: int = 13, autoreg: bool = False):
    """Seeded synthetic conditioning with the shapes the layout pipeline produces:
    scaled_gt_boxes [B,13,9], gt_boxes_2d [B,13,4], is_valid_obj [B,13], concat_cond [B,10,H,W]
    (one-hot class map (9) + normalised centre depth), optional autoregressive_cond [B,1,H,W]."""
    g = torch.Generator().manual_seed(1000 + seed)
    boxes = torch.rand(B, n_obj, 9, generator=g)
    boxes[..., 8] = torch.randint(0, 9, (B, n_obj), generator=g).float()
    b2 = torch.rand(B, n_obj, 4, generator=g)
    b2 = torch.cat([b2[..., :2].min(b2[..., 2:]), b2[..., :2].max(b2[..., 2:])], -1)
    valid = (torch.rand(B, n_obj, generator=g) > 0.2).float()
    cls_map = torch.zeros(B, H, W, dtype=torch.long)
    depth = torch.zeros(B, 1, H, W)
    for b in range(B):
        for o in range(n_obj):
            x1, y1, x2, y2 = (b2[b, o] * torch.tensor([W, H, W, H])).long().tolist()
            cls_map[b, y1:y2, x1:x2] = int(boxes[b, o, 8])
            depth[b, 0, y1:y2, x1:x2] = float(torch.rand(1, generator=g))
    concat = torch.cat([F.one_hot(cls_map, 9).permute(0, 3, 1, 2).float(), depth], dim=1)
    batch = dict(scaled_gt_boxes=boxes, gt_boxes_2d=b2, is_valid_obj=valid, concat_cond=concat)
    if autoreg:
        batch["autoregressive_cond"] = torch.rand(B, 1, H, W, generator=g) * 2 - 1
    return batch
