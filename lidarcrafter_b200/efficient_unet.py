"""B200-native drop-in for the reference ``EfficientUNet`` (lidargen/models/unets/efficient_unet.py:193-300).

Same constructor kwargs, same ``state_dict`` key names / shapes, same ``forward(images, timesteps)``
contract (NCHW fp32 in, NCHW fp32 out) and the attributes the diffusion wrapper reads
(``in_channels``, ``resolution``, ``coords``).  The parameters live in ordinary ``nn`` containers so
reference checkpoints load unchanged; ``forward`` never calls a PyTorch operator on the hot path -- it
replays a static plan of libb200lidar kernels (engine.py).  No CPU / eager fallback exists.
"""
from __future__ import annotations

import math
from typing import Iterable

import torch
from torch import nn

from . import _lib
from .engine import Act, Plan, PlanBuilder, _ptr, _sp, precision_parts


def _n_tuple(x, n):
    if isinstance(x, Iterable):
        x = tuple(x)
        assert len(x) == n
        return x
    return (x,) * n


def generate_polar_coords(H: int, W: int) -> torch.Tensor:
    """encoding.py:80-89 (default coords buffer; the builders overwrite it with real ray angles)."""
    phi = (0.5 - torch.arange(H) / H) * torch.pi
    theta = (1 - torch.arange(W) / W) * 2 * torch.pi - torch.pi
    phi, theta = torch.meshgrid([phi, theta], indexing="ij")
    return torch.stack([phi, theta])[None]


def fourier_features(coords: torch.Tensor, resolution) -> torch.Tensor:
    """encoding.py:120-149: [1,2,H,W] angles -> [1, 2(L_h+L_w), H, W] (sin block then cos block)."""
    L_h = int(math.ceil(math.log2(resolution[0])))
    L_w = int(math.ceil(math.log2(resolution[1])))
    fh = torch.cat([torch.arange(L_h).float().exp2(), torch.zeros(L_w)])
    fw = torch.cat([torch.zeros(L_h), torch.arange(L_w).float().exp2()])
    ang = coords[:, 0:1] * fh[None, :, None, None] + coords[:, 1:2] * fw[None, :, None, None]
    return torch.cat([ang.sin(), ang.cos()], dim=1)


# ---------------------------------------------------------------------------------------------
# parameter containers (names mirror the reference so state_dicts are interchangeable)
# ---------------------------------------------------------------------------------------------
class _ConvP(nn.Conv2d):
    """weight/bias holder for ops.Conv2d (ops.py:149-173); never executed."""

    def __init__(self, cin, cout, k):
        super().__init__(cin, cout, k, 1, 0, bias=True)

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container only; the CUDA plan executes this layer")


class _KernelBuf(nn.Module):
    """ops.Resample's registered FIR ``kernel`` buffer (ops.py:91-96)."""

    def __init__(self, up: bool):
        super().__init__()
        k = torch.tensor([1.0, 3.0, 3.0, 1.0])
        k = k / k.sum()
        if up:
            k = k * 2.0
        self.register_buffer("kernel", k)


class _AdaGNP(nn.Module):
    def __init__(self, emb, cout):
        super().__init__()
        self.proj = nn.Sequential(nn.Identity(), nn.Linear(emb, cout * 2), nn.Identity())


def _zero(m: nn.Module):
    for p in m.parameters():
        p.data.zero_()


class _ResBlockP(nn.Module):
    def __init__(self, cin, cout, emb, groups, eps):
        super().__init__()
        self.norm1 = nn.GroupNorm(groups, cin, eps)
        self.conv1 = _ConvP(cin, cout, 3)
        self.norm2 = _AdaGNP(emb, cout)
        self.conv2 = _ConvP(cout, cout, 3)
        _zero(self.conv2)
        self.skip = _ConvP(cin, cout, 1) if cin != cout else nn.Identity()
        self.register_buffer("scale", torch.tensor(1 / math.sqrt(2)).float())
        self.cin, self.cout = cin, cout


class _MHAP(nn.Module):
    def __init__(self, E):
        super().__init__()
        self.in_proj_weight = nn.Parameter(torch.empty(3 * E, E))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * E))
        self.out_proj = nn.Linear(E, E)
        nn.init.xavier_uniform_(self.in_proj_weight)


class _SelfAttnP(nn.Module):
    def __init__(self, C, heads, groups, eps):
        super().__init__()
        self.norm = nn.GroupNorm(groups, C, eps)
        self.attn = _MHAP(C)
        _zero(self.attn.out_proj)
        self.register_buffer("scale", torch.tensor(1 / math.sqrt(2)).float())
        self.heads = heads


class _BlockP(nn.Module):
    def __init__(self, cin, cout, n_res, emb, groups, eps, heads, attn=False, up=1, down=1):
        super().__init__()
        self.downsample = (nn.Sequential(_ConvP(cin, cout, 3), _KernelBuf(False)) if down > 1 else nn.Identity())
        self.residual_blocks = nn.Sequential()
        for i in range(n_res):
            self.residual_blocks.append(
                _ResBlockP(cout if (i != 0 or down > 1) else cin, cout, emb, groups, eps))
        self.self_attn_block = _SelfAttnP(cout, heads, groups, eps) if attn else nn.Identity()
        self.upsample = (nn.Sequential(_KernelBuf(True), _ConvP(cout, cout, 3)) if up > 1 else nn.Identity())
        self.has_down, self.has_up, self.has_attn = down > 1, up > 1, attn


class _FourierP(nn.Module):
    def __init__(self, resolution):
        super().__init__()
        L_h = int(math.ceil(math.log2(resolution[0])))
        L_w = int(math.ceil(math.log2(resolution[1])))
        fh = torch.cat([torch.arange(L_h).float().exp2(), torch.zeros(L_w)])
        fw = torch.cat([torch.zeros(L_h), torch.arange(L_w).float().exp2()])
        self.register_buffer("freqs", torch.stack([fh, fw], dim=-1)[..., None, None])
        self.register_buffer("phase", torch.zeros(L_h + L_w))
        self.extra_ch = 2 * (L_h + L_w)


class EfficientUNet(nn.Module):
    """Registry key ``"efficient_unet"`` (lidargen/models/unets/__init__.py:22-37)."""

    def __init__(self, in_channels: int, resolution, out_channels: int | None = None, base_channels: int = 128,
                 temb_channels: int | None = None, channel_multiplier=(1, 2, 4, 8), num_residual_blocks=(3, 3, 3, 3),
                 gn_num_groups: int = 32 // 4, gn_eps: float = 1e-6, attn_num_heads: int = 8,
                 coords_encoding: str | None = "spherical_harmonics", ring: bool = True):
        super().__init__()
        self.resolution = _n_tuple(resolution, 2)
        self.in_channels = in_channels
        self.out_channels = in_channels if out_channels is None else out_channels
        temb_channels = base_channels * 4 if temb_channels is None else temb_channels
        self.base_channels, self.temb_channels = base_channels, temb_channels
        self.gn_num_groups, self.gn_eps, self.ring = gn_num_groups, gn_eps, ring
        self.attn_num_heads = attn_num_heads

        self.register_buffer("coords", generate_polar_coords(*self.resolution))
        if coords_encoding == "fourier_features":
            self.coords_encoding = _FourierP(self.resolution)
            cin0 = in_channels + self.coords_encoding.extra_ch
        elif coords_encoding is None:
            self.coords_encoding = None
            cin0 = in_channels
        else:
            raise NotImplementedError(
                f"coords_encoding={coords_encoding!r}: the nuScenes hot path uses 'fourier_features' "
                "(lidargen/utils/configs/option_unet_nusc.py:19)")

        self.time_embedding = nn.Sequential(nn.Identity(), nn.Linear(base_channels, temb_channels), nn.Identity(),
                                            nn.Linear(temb_channels, temb_channels))
        mult = _n_tuple(channel_multiplier, 4)
        C = [base_channels] + [base_channels * m for m in mult]
        N = _n_tuple(num_residual_blocks, 4)
        kw = dict(emb=temb_channels, groups=gn_num_groups, eps=gn_eps, heads=attn_num_heads)
        self.in_conv = _ConvP(cin0, C[0], 3)
        self.d_block1 = _BlockP(C[0], C[1], N[0], **kw)
        self.d_block2 = _BlockP(C[1], C[2], N[1], down=2, **kw)
        self.d_block3 = _BlockP(C[2], C[3], N[2], down=2, **kw)
        self.d_block4 = _BlockP(C[3], C[4], N[3], down=2, attn=True, **kw)
        self.u_block4 = _BlockP(C[4], C[3], N[3], up=2, attn=True, **kw)
        self.u_block3 = _BlockP(C[3] + C[3], C[2], N[2], up=2, **kw)
        self.u_block2 = _BlockP(C[2] + C[2], C[1], N[1], up=2, **kw)
        self.u_block1 = _BlockP(C[1] + C[1], C[0], N[0], **kw)
        self.out_conv = _ConvP(C[0], self.out_channels, 3)
        _zero(self.out_conv)

        self._plans: dict = {}
        self.conv_impl = "tc"   # "ffma" = CUDA-core cross-check path (tests only)
        # "fp16f8" (default): fp16 MMA + ONE e4m3 MMA for both correction cross terms per product: 5e-5 relative per forward,
        #           1.3e-4 at the end of a 50-step DDIM trajectory (tolerance 1e-3; measured: tests/test_gpu_unet.py
        #           the 50-step trajectory test, profiles/r02_trajectory_*.json) at 2/3 of the tensor work
        # "fp16x3": error-compensated fp16 split (3 MMAs / product, fp32-grade: 2e-6 per forward, 2e-5 after 50 steps)
        # "fp16"  : one MMA / product (~2e-3 relative through the UNet: outside the tolerance, measurement only)
        self.precision = "fp16f8"
        self.register_load_state_dict_post_hook(lambda mod, keys: mod.invalidate())

    # ------------------------------------------------------------------------------------------
    def invalidate(self):
        """Drop cached plans (weights are repacked on the next forward)."""
        self._plans = {}

    def _apply(self, fn, *a, **k):
        self._plans = {}
        return super()._apply(fn, *a, **k)

    def __setattr__(self, name, value):
        # the builders assign ``model.coords = get_linear_ray_angles(...)`` after construction
        # (lidargen/utils/inference.py:281-282): cached constants depend on it
        if name == "coords" and "_plans" in self.__dict__:
            self.__dict__["_plans"] = {}
        super().__setattr__(name, value)

    def get_plan(self, B: int) -> "EfficientUNetPlan":
        parts = precision_parts(self.precision)
        key = (B, self.conv_impl, self.precision)
        if key not in self._plans:
            self._plans[key] = EfficientUNetPlan(self, B, self.conv_impl, parts)
        return self._plans[key]

    @torch.no_grad()
    def forward(self, images: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        if not images.is_cuda and _lib._TEST_LIB is None:
            raise _lib.B200LidarError("lidarcrafter_b200.EfficientUNet runs on a B200 only (no CPU fallback)")
        B = images.shape[0]
        if timesteps.dim() == 0:
            timesteps = timesteps[None].repeat_interleave(B, dim=0)
        plan = self.get_plan(B)
        return plan(images, timesteps)


# ---------------------------------------------------------------------------------------------
# the plan
# ---------------------------------------------------------------------------------------------
class EfficientUNetPlan:
    """Static launch list of one EfficientUNet forward at batch B (efficient_unet.py:274-300)."""

    def __init__(self, m: EfficientUNet, B: int, conv_impl: str = "tc", parts: int = 2):
        dev = m.in_conv.weight.device
        self.lib = _lib.get_lib()
        _lib.require_b200(dev.index or 0)
        self.m, self.B, self.dev = m, B, dev
        H, W = m.resolution
        self.H, self.W = H, W
        stream = _lib.current_stream(dev)
        plan = Plan(self.lib, dev, B, conv_impl, parts)
        pb = PlanBuilder(plan, m.ring, stream)
        self.plan, self.pb = plan, pb
        G, eps = m.gn_num_groups, m.gn_eps
        E = m.temb_channels

        # static I/O buffers (reference layout: NCHW fp32)
        self.x_in = plan.f32(B, m.in_channels, H, W)
        self.t_in = plan.f32(B)
        self.pred = plan.f32(B, m.out_channels, H, W)

        # ---- K4: time embedding + all AdaGN projections in one launch ----
        blocks = [m.d_block1, m.d_block2, m.d_block3, m.d_block4, m.u_block4, m.u_block3, m.u_block2, m.u_block1]
        wp, bp, self.ada_off = [], [], {}
        off = 0
        for blk in blocks:
            for rb in blk.residual_blocks:
                lin = rb.norm2.proj[1]
                wp.append(lin.weight.detach().float())
                bp.append(lin.bias.detach().float())
                self.ada_off[id(rb)] = off
                off += lin.weight.shape[0]
        self.P = off
        self.wp = torch.cat(wp, 0).contiguous()
        self.bp = torch.cat(bp, 0).contiguous()
        self.temb = plan.f32(B, E)
        self.ada = plan.f32(B, self.P)
        te = m.time_embedding
        tw = [te[1].weight.detach().float().contiguous(), te[1].bias.detach().float().contiguous(),
              te[3].weight.detach().float().contiguous(), te[3].bias.detach().float().contiguous()]
        plan.bufs += tw
        plan.add(self.lib.time_embed, _ptr(self.t_in), _ptr(tw[0]), _ptr(tw[1]), _ptr(tw[2]), _ptr(tw[3]), 0,
                 _ptr(self.wp), _ptr(self.bp), _ptr(self.temb), _ptr(self.ada), B, m.base_channels, E, self.P,
                 name="time_embed", nbytes=4.0 * self.P * E)

        # ---- in_conv: constant (Fourier) part folded once, dynamic 2-channel part per step ----
        w_in = m.in_conv.weight.detach().float()
        cx = m.in_channels
        assert cx <= 4, "in_conv dynamic part supports <= 4 channels"
        C0 = w_in.shape[0]
        self.cst = plan.f32(1, H * W, C0)
        if m.coords_encoding is not None:
            ff = fourier_features(m.coords.detach().float().cpu(), (H, W))          # [1, 30, H, W] (constant table)
            ff_nhwc = ff.permute(0, 2, 3, 1).contiguous().to(dev)
            w_c = w_in[:, cx:].contiguous()
            b_in = m.in_conv.bias.detach().float().contiguous()
            self.lib.conv_direct_f32(_ptr(ff_nhwc), _ptr(w_c), _ptr(b_in), _ptr(self.cst), 1, H, W, w_c.shape[1], C0,
                                     3, 1 if m.ring else 0, stream)
            plan.bufs += [ff_nhwc, w_c, b_in]
        else:
            self.cst.copy_(m.in_conv.bias.detach().float()[None, None, :].expand(1, H * W, C0))
        w_dyn = w_in[:, :cx].contiguous()
        plan.bufs.append(w_dyn)
        h0 = plan.f32(B, H * W, C0)
        st0 = plan.new_stats(C0)
        plan.add(self.lib.in_conv, _ptr(self.x_in), _ptr(w_dyn), _ptr(self.cst), 0, _ptr(h0), _sp(st0), B, H, W, cx, C0,
                 1 if m.ring else 0, name="in_conv", nbytes=4.0 * H * W * C0 * (B + 1))
        h = Act(h0, H, W, C0, st0)

        # ---- U-Net ----
        h1 = self._block(m.d_block1, [h])
        h2 = self._block(m.d_block2, [h1])
        h3 = self._block(m.d_block3, [h2])
        h4 = self._block(m.d_block4, [h3])
        u = self._block(m.u_block4, [h4])
        u = self._block(m.u_block3, [u, h3])
        u = self._block(m.u_block2, [u, h2])
        u = self._block(m.u_block1, [u, h1])

        # ---- out_conv (raw fp32 activations, no norm: efficient_unet.py:298) ----
        w_out = m.out_conv.weight.detach().float().contiguous()
        b_out = m.out_conv.bias.detach().float().contiguous()
        plan.bufs += [w_out, b_out]
        plan.add(self.lib.out_conv, _ptr(u.t), 0, _ptr(w_out), _ptr(b_out), _ptr(self.pred), B, u.H, u.W, u.C,
                 m.out_channels, 1 if m.ring else 0, name="out_conv", nbytes=4.0 * B * u.H * u.W * (u.C + m.out_channels))
        plan.finalize()
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)

    # ------------------------------------------------------------------------------------------
    def _resblock(self, rb: _ResBlockP, srcs: list[Act]) -> Act:
        """efficient_unet.py:61-115."""
        pb, m = self.pb, self.m
        G, eps = m.gn_num_groups, m.gn_eps
        H, W = srcs[0].H, srcs[0].W
        has_skip = not isinstance(rb.skip, nn.Identity)
        cin = sum(a.C for a in srcs)
        self._n_rb = getattr(self, "_n_rb", 0) + 1
        self.plan.tag = f"resblock {H}x{W} {cin}->{rb.cout}{' skip' if has_skip else ''} #{self._n_rb}"
        try:
            return self._resblock_ops(rb, srcs, has_skip)
        finally:
            self.plan.tag = ""

    def _resblock_ops(self, rb: _ResBlockP, srcs: list[Act], has_skip: bool) -> Act:
        pb, m = self.pb, self.m
        G, eps = m.gn_num_groups, m.gn_eps
        H, W = srcs[0].H, srcs[0].W
        # conv1(silu(norm1(x))): GroupNorm-apply + SiLU + operand split run inside the conv launch (b200_conv_gn_tc)
        hmid, st_h = pb.conv_gn(srcs, rb.conv1.weight, rb.conv1.bias, None, 1.0, True, gamma=rb.norm1.weight,
                                beta=rb.norm1.bias, groups=G, eps=eps, silu=True)
        if not has_skip:
            assert len(srcs) == 1
            res = srcs[0].t
        else:           # 1x1 skip conv on the raw (un-normalised) input
            res, _ = pb.conv_gn(srcs, rb.skip.weight, rb.skip.bias, None, 1.0, False, normalize=False)
        # conv2(silu(AdaGN(h, temb))) + skip, * 1/sqrt(2)
        out, st = pb.conv_gn([Act(hmid, H, W, rb.cout, st_h)], rb.conv2.weight, rb.conv2.bias, res, float(rb.scale), True,
                             groups=G, eps=eps, silu=True, ada=self.ada, ada_stride=self.P, ada_off=self.ada_off[id(rb)])
        return Act(out, H, W, rb.cout, st)

    def _attention(self, ab: _SelfAttnP, x: Act) -> Act:
        """efficient_unet.py:28-58."""
        pb, m = self.pb, self.m
        E, nh = x.C, ab.heads
        T = x.H * x.W
        w_qkv = ab.attn.in_proj_weight.detach().reshape(3 * E, E, 1, 1)
        qkv, _ = pb.conv_gn([x], w_qkv, ab.attn.in_proj_bias, None, 1.0, False, gamma=ab.norm.weight, beta=ab.norm.bias,
                            groups=m.gn_num_groups, eps=m.gn_eps, silu=False)
        att = self.plan.operand(x.H, x.W, E)
        d = E // nh
        ws = torch.empty(max(int(self.lib.flash_attention_workspace(self.B, nh, T, 0, d, d)), 16), dtype=torch.uint8, device=self.dev)
        self.plan.bufs.append(ws)           # packed Q / K / V^T tile images of the tcgen05 attention
        self.plan.add(self.lib.flash_attention, _ptr(qkv), E, _ptr(att), x.W, self.plan.parts, self.B, nh, T,
                      1.0 / math.sqrt(d), _ptr(ws), name="attention", flops=4.0 * self.B * nh * T * T * d)
        self.plan.flops += 4.0 * self.B * nh * T * T * d
        w_o = ab.attn.out_proj.weight.detach().reshape(E, E, 1, 1)
        out, st = pb.conv(att, x.H, x.W, w_o, ab.attn.out_proj.bias, x.t, float(ab.scale), True)
        return Act(out, x.H, x.W, E, st)

    def _block(self, blk: _BlockP, srcs: list[Act]) -> Act:
        """efficient_unet.py:118-190."""
        pb = self.pb
        if blk.has_down:
            assert len(srcs) == 1
            x = srcs[0]
            conv = blk.downsample[0]
            y, _ = pb.conv_gn([x], conv.weight, conv.bias, None, 1.0, False, normalize=False)
            h = pb.fir(Act(y, x.H, x.W, conv.weight.shape[0]), up=False, want_stats=True)
            srcs = [h]
        for rb in blk.residual_blocks:
            srcs = [self._resblock(rb, srcs)]
        h = srcs[0]
        if blk.has_attn:
            h = self._attention(blk.self_attn_block, h)
        if blk.has_up:
            conv = blk.upsample[1]
            x16, Hu, Wu = pb.fir_up_operand(h)      # Resample(up) -> conv operand in one pass (no fp32 round trip)
            y, st = pb.conv(x16, Hu, Wu, conv.weight, conv.bias, None, 1.0, True)
            h = Act(y, Hu, Wu, conv.weight.shape[0], st)
        return h

    # ------------------------------------------------------------------------------------------
    def launch(self, stream: int | None = None):
        """Run the plan on the static buffers (x_in, t_in -> pred)."""
        if stream is None:
            stream = _lib.current_stream(self.dev)
        self.plan.run(stream)

    def __call__(self, images: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        self.x_in.copy_(images)
        self.t_in.copy_(timesteps)
        self.launch()
        return self.pred.clone()
