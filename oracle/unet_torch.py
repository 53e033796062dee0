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
        if "norm" in k and leaf == "weight" and v.dim() == 1:
            out[k] = 1 + 0.1 * torch.randn(v.shape, generator=g)
        elif "norm" in k and leaf == "bias" and "proj" not in k:
            out[k] = 0.1 * torch.randn(v.shape, generator=g)
        elif leaf in ("bias", "in_proj_bias") or v.dim() == 1:
            out[k] = 0.02 * torch.randn(v.shape, generator=g)
        else:
            fan_in = v[0].numel()
            out[k] = torch.randn(v.shape, generator=g) / math.sqrt(fan_in)
    return out
