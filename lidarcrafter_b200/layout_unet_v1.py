"""B200-native drop-in for the reference ``LayoutUnetV1`` (lidargen/models/unets/layout_unet_v1.py:599-902):
the layout-conditioned range-image denoiser (scale-shift ResBlocks with FIR up/down, ObjectAwareCrossAttention
at 1/4 and 1/8 resolution, concat-conditioning, Fourier coordinate channels).

Same constructor kwargs (nuscenes-box-layout-v2..v6 / nuscenes-auto-reg* configs), same ``state_dict`` keys
(505 tensors, 70 105 602 parameters for the v3 config) and the same call
``model(x, {"time_condition": log_snr[B], "other_condition": layout_encoder_output_dict})``.
Everything that depends only on the condition (it is fixed for a whole ``sample()``) is folded once in
``LayoutUnetPlan.set_condition``: the conv of [concat_cond, Fourier] channels of the first conv, and per attention
layer the normalised positional embeddings and the layout key/value projections.  Per step the plan launches only
libb200lidar kernels.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F
from torch import nn

from . import _lib
from .efficient_unet import _ConvP, _FourierP, _KernelBuf, _n_tuple, _zero, fourier_features, generate_polar_coords
from .engine import Act, Plan, PlanBuilder, _ptr, _sp, precision_parts

GN_GROUPS = 32
GN_EPS = 1e-5


class _LResBlockP(nn.Module):
    """parameters of layout_unet_v1.ResBlock (:143-249, use_scale_shift_norm=True)"""

    def __init__(self, cin, cout, emb, up=False, down=False):
        super().__init__()
        self.in_layers = nn.Sequential(nn.GroupNorm(GN_GROUPS, cin), nn.Identity(), _ConvP(cin, cout, 3))
        self.op = _KernelBuf(up) if (up or down) else nn.Identity()
        self.emb_layers = nn.Sequential(nn.Identity(), nn.Linear(emb, 2 * cout))
        self.out_layers = nn.Sequential(nn.GroupNorm(GN_GROUPS, cout), nn.Identity(), nn.Identity(), _ConvP(cout, cout, 3))
        _zero(self.out_layers[3])
        self.skip_connection = nn.Identity() if cin == cout else _ConvP(cin, cout, 1)
        self.cin, self.cout, self.up, self.down = cin, cout, up, down


class _OAAttnP(nn.Module):
    """parameters of ObjectAwareCrossAttention (:347-414), norm_first=False, positional scale 1.0"""

    def __init__(self, C, head_ch, enc):
        super().__init__()
        self.qkv_projector = nn.Conv1d(C, 3 * C, 1)
        self.norm_for_qkv = nn.GroupNorm(GN_GROUPS, C)
        self.layout_content_embedding_projector = nn.Conv1d(enc, 2 * C, 1)
        self.layout_position_embedding_projector = nn.Conv1d(enc, C, 1)
        self.norm_for_obj_class_embedding = nn.GroupNorm(GN_GROUPS, enc)
        self.norm_for_layout_positional_embedding = nn.GroupNorm(GN_GROUPS, C)
        self.norm_for_image_patch_positional_embedding = nn.GroupNorm(GN_GROUPS, C)
        self.proj_out = nn.Conv1d(C, C, 1)
        _zero(self.proj_out)
        self.channels, self.num_heads = C, C // head_ch


class LayoutUnetV1(nn.Module):
    """Registry key ``"layout_unet_v1"``."""

    def __init__(self, in_channels, resolution, model_channels, out_channels, num_res_blocks, attention_ds,
                 encoder_channels=None, dropout=0, channel_mult=(1, 2, 4, 8), conv_resample=True, dims=2,
                 use_checkpoint=False, use_fp16=False, num_heads=1, num_head_channels=-1, num_heads_upsample=-1,
                 use_scale_shift_norm=False, resblock_updown=False, use_positional_embedding_for_attention=False,
                 image_size=256, attention_block_type="GLIDE", num_attention_blocks=1, use_key_padding_mask=False,
                 channels_scale_for_positional_embedding=1.0, norm_first=False, norm_for_obj_embedding=False,
                 coords_encoding="fourier_features", **kwargs):
        super().__init__()
        unsupported = []
        if attention_block_type != "ObjectAwareCrossAttention": unsupported.append("attention_block_type")
        if not use_scale_shift_norm: unsupported.append("use_scale_shift_norm=False")
        if not resblock_updown: unsupported.append("resblock_updown=False")
        if use_fp16: unsupported.append("use_fp16")
        if norm_first or norm_for_obj_embedding or use_key_padding_mask: unsupported.append("norm_first/key_padding_mask")
        if channels_scale_for_positional_embedding != 1.0: unsupported.append("channels_scale_for_positional_embedding")
        if num_attention_blocks != 1 or num_head_channels != 32 or dims != 2: unsupported.append("attention geometry")
        if coords_encoding != "fourier_features": unsupported.append("coords_encoding")
        if unsupported:
            raise NotImplementedError("LayoutUnetV1 on B200 implements the nuScenes layout configs "
                                      "(lidargen/utils/configs/option_nusc_box_layout_v3.py:10-33); unsupported: "
                                      + ", ".join(unsupported))
        self.in_channels = in_channels
        self.resolution = _n_tuple(resolution, 2)
        self.model_channels, self.out_channels = model_channels, out_channels
        self.num_res_blocks, self.attention_ds = num_res_blocks, list(attention_ds)
        self.channel_mult, self.encoder_channels = list(channel_mult), encoder_channels
        self.num_head_channels, self.image_size = num_head_channels, image_size
        self.register_buffer("coords", generate_polar_coords(*self.resolution))
        self.coords_encoding = _FourierP(self.resolution)
        cin0 = in_channels + self.coords_encoding.extra_ch
        E = model_channels * 4
        self.time_embed = nn.Sequential(nn.Identity(), nn.Linear(model_channels, E), nn.Identity(), nn.Linear(E, E))

        def attn(ch):
            return _OAAttnP(ch, num_head_channels, encoder_channels)
        ch = int(channel_mult[0] * model_channels)
        self.input_blocks = nn.ModuleList([nn.Sequential(_ConvP(cin0, ch, 3))])
        chans, ds = [ch], 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [_LResBlockP(ch, int(mult * model_channels), E)]
                ch = int(mult * model_channels)
                if ds in self.attention_ds:
                    layers.append(attn(ch))
                self.input_blocks.append(nn.Sequential(*layers))
                chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(nn.Sequential(_LResBlockP(ch, ch, E, down=True)))
                chans.append(ch)
                ds *= 2
        self.middle_block = nn.Sequential(_LResBlockP(ch, ch, E), attn(ch), _LResBlockP(ch, ch, E))
        self.output_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                ich = chans.pop()
                layers = [_LResBlockP(ch + ich, int(model_channels * mult), E)]
                ch = int(model_channels * mult)
                if ds in self.attention_ds:
                    layers.append(attn(ch))
                if level and i == num_res_blocks:
                    layers.append(_LResBlockP(ch, ch, E, up=True))
                    ds //= 2
                self.output_blocks.append(nn.Sequential(*layers))
        self.out = nn.Sequential(nn.GroupNorm(GN_GROUPS, ch), nn.Identity(), _ConvP(ch, out_channels, 3))
        _zero(self.out[2])
        self._plans: dict = {}
        self.conv_impl = "tc"
        self.precision = "fp16x3"
        self.register_load_state_dict_post_hook(lambda mod, keys: mod.invalidate())

    def invalidate(self):
        self._plans = {}

    def _apply(self, fn, *a, **k):
        self._plans = {}
        return super()._apply(fn, *a, **k)

    def __setattr__(self, name, value):
        if name == "coords" and "_plans" in self.__dict__:
            self.__dict__["_plans"] = {}
        super().__setattr__(name, value)

    def get_plan(self, B: int) -> "LayoutUnetPlan":
        parts = precision_parts(self.precision)
        key = (B, self.conv_impl, self.precision)
        if key not in self._plans:
            self._plans[key] = LayoutUnetPlan(self, B, self.conv_impl, parts)
        return self._plans[key]

    @torch.no_grad()
    def forward(self, x: torch.Tensor, cond_dict: dict) -> torch.Tensor:
        if not x.is_cuda and _lib._TEST_LIB is None:
            raise _lib.B200LidarError("lidarcrafter_b200.LayoutUnetV1 runs on a B200 only (no CPU fallback)")
        plan = self.get_plan(x.shape[0])
        plan.set_condition(cond_dict["other_condition"])
        return plan(x, cond_dict["time_condition"])


class LayoutUnetPlan:
    """Static launch list of one LayoutUnetV1 forward at batch B (layout_unet_v1.py:866-902)."""

    def __init__(self, m: LayoutUnetV1, B: int, conv_impl: str = "tc", parts: int = 2):
        dev = m.out[2].weight.device
        self.lib = _lib.get_lib()
        _lib.require_b200(dev.index or 0)
        self.m, self.B, self.dev = m, B, dev
        H, W = m.resolution
        self.H, self.W = H, W
        self.stream = _lib.current_stream(dev)
        plan = Plan(self.lib, dev, B, conv_impl, parts)
        pb = PlanBuilder(plan, True, self.stream)
        self.plan, self.pb = plan, pb
        E = m.model_channels * 4

        self.x_in = plan.f32(B, m.out_channels, H, W)      # the dynamic (noisy) channels
        self.t_in = plan.f32(B)
        self.pred = plan.f32(B, m.out_channels, H, W)
        self.xf_proj = plan.f32(B, E)

        # ---- time embedding (+ xf_proj) and the stacked scale/shift projections of all ResBlocks ----
        self.resblocks = [l for blk in list(m.input_blocks) + [m.middle_block] + list(m.output_blocks) for l in blk
                          if isinstance(l, _LResBlockP)]
        wp, bp, self.ada_off, off = [], [], {}, 0
        for rb in self.resblocks:
            lin = rb.emb_layers[1]
            wp.append(lin.weight.detach().float()); bp.append(lin.bias.detach().float())
            self.ada_off[id(rb)] = off
            off += lin.weight.shape[0]
        self.P = off
        self.wp, self.bp = torch.cat(wp, 0).contiguous(), torch.cat(bp, 0).contiguous()
        self.temb, self.ada = plan.f32(B, E), plan.f32(B, self.P)
        te = m.time_embed
        tw = [te[1].weight.detach().float().contiguous(), te[1].bias.detach().float().contiguous(),
              te[3].weight.detach().float().contiguous(), te[3].bias.detach().float().contiguous()]
        plan.bufs += tw
        plan.add(self.lib.time_embed, _ptr(self.t_in), _ptr(tw[0]), _ptr(tw[1]), _ptr(tw[2]), _ptr(tw[3]),
                 _ptr(self.xf_proj), _ptr(self.wp), _ptr(self.bp), _ptr(self.temb), _ptr(self.ada), B, m.model_channels, E,
                 self.P, name="time_embed", nbytes=4.0 * self.P * E)

        # ---- first conv: per-sample constant part (concat_cond + Fourier channels) folded in set_condition ----
        conv0 = m.input_blocks[0][0]
        w0 = conv0.weight.detach().float()
        C0 = w0.shape[0]
        self.cx = m.out_channels
        self.n_cond = m.in_channels - self.cx
        self.w_dyn = w0[:, :self.cx].contiguous()
        self.w_cst = w0[:, self.cx:].contiguous()
        self.b0 = conv0.bias.detach().float().contiguous()
        self.cst_in = plan.f32(B, H * W, self.n_cond + m.coords_encoding.extra_ch)
        self.cst = plan.f32(B, H * W, C0)
        ff = fourier_features(m.coords.detach().float().cpu(), (H, W)).permute(0, 2, 3, 1).reshape(1, H * W, -1)
        self.cst_in[:, :, self.n_cond:] = ff.to(dev)
        h0 = plan.f32(B, H * W, C0)
        st0 = plan.new_stats(C0)
        plan.add(self.lib.in_conv, _ptr(self.x_in), _ptr(self.w_dyn), _ptr(self.cst), 1, _ptr(h0), _sp(st0), B, H, W,
                 self.cx, C0, 1, name="in_conv", nbytes=4.0 * H * W * C0 * 2 * B)
        h = Act(h0, H, W, C0, st0)

        # ---- U-Net body ----
        self.attn_consts = []      # (module, resolution key, buffers) filled by set_condition
        hs = [h]
        for blk in list(m.input_blocks)[1:]:
            h = self._run(blk, [h])
            hs.append(h)
        h = self._run(m.middle_block, [h])
        for blk in m.output_blocks:
            h = self._run(blk, [h, hs.pop()])

        # ---- out: GN -> SiLU -> ring conv 64 -> 2 ----
        g = m.out[0].weight.detach().float().contiguous()
        b = m.out[0].bias.detach().float().contiguous()
        a_out = plan.f32(B, h.H * h.W, h.C)
        plan.bufs += [g, b]
        plan.add(self.lib.gn_act_f32, _ptr(h.t), _sp(h.stats), _ptr(g), _ptr(b), GN_GROUPS, GN_EPS, 1, _ptr(a_out), B,
                 h.H * h.W, h.C, name="gn_act_f32", nbytes=8.0 * B * h.H * h.W * h.C)
        w_out = m.out[2].weight.detach().float().contiguous()
        b_out = m.out[2].bias.detach().float().contiguous()
        plan.bufs += [w_out, b_out]
        plan.add(self.lib.out_conv, _ptr(a_out), 0, _ptr(w_out), _ptr(b_out), _ptr(self.pred), B, h.H, h.W, h.C,
                 m.out_channels, 1, name="out_conv", nbytes=4.0 * B * h.H * h.W * (h.C + m.out_channels))
        plan.finalize()
        if dev.type == "cuda":
            torch.cuda.synchronize(dev)

    # ------------------------------------------------------------------------------------------
    def _run(self, blk, srcs):
        for layer in blk:
            if isinstance(layer, _LResBlockP):
                srcs = [self._resblock(layer, srcs)]
            elif isinstance(layer, _OAAttnP):
                srcs = [self._attention(layer, srcs[0])]
            else:
                raise TypeError(type(layer))
        return srcs[0]

    def _resblock(self, rb: _LResBlockP, srcs):
        """layout_unet_v1.py:222-249."""
        pb, plan = self.pb, self.plan
        n1, conv1, n2, conv2 = rb.in_layers[0], rb.in_layers[2], rb.out_layers[0], rb.out_layers[3]
        x0 = srcs[0]
        H, W = x0.H, x0.W
        if rb.up or rb.down:
            assert len(srcs) == 1
            g = n1.weight.detach().float().contiguous(); b = n1.bias.detach().float().contiguous()
            t = plan.f32(self.B, H * W, x0.C)
            plan.bufs += [g, b]
            plan.add(self.lib.gn_act_f32, _ptr(x0.t), _sp(x0.stats), _ptr(g), _ptr(b), GN_GROUPS, GN_EPS, 1, _ptr(t),
                     self.B, H * W, x0.C, name="gn_act_f32", nbytes=8.0 * self.B * H * W * x0.C)
            xr = pb.fir(x0, up=rb.up, want_stats=False)
            if rb.up:       # Resample(up) written straight into conv1's operand
                a1, H, W = pb.fir_up_operand(Act(t, H, W, x0.C))
                hmid, st_h = pb.conv(a1, H, W, conv1.weight, conv1.bias, None, 1.0, True)
            else:           # Resample(down) in fp32, converted to the operand inside the conv launch
                tr = pb.fir(Act(t, H, W, x0.C), up=False, want_stats=False)
                H, W = tr.H, tr.W
                hmid, st_h = pb.conv_gn([tr], conv1.weight, conv1.bias, None, 1.0, True, normalize=False)
            res = xr.t
        else:
            # conv1(silu(GN32(x))) with the normalisation fused in front of the conv (b200_conv_gn_tc)
            hmid, st_h = pb.conv_gn(srcs, conv1.weight, conv1.bias, None, 1.0, True, gamma=n1.weight, beta=n1.bias,
                                    groups=GN_GROUPS, eps=GN_EPS, silu=True)
            if isinstance(rb.skip_connection, nn.Identity):
                assert len(srcs) == 1
                res = x0.t
            else:
                res, _ = pb.conv_gn(srcs, rb.skip_connection.weight, rb.skip_connection.bias, None, 1.0, False,
                                    normalize=False)
        # conv2(silu(GN32(h) * (1 + scale) + shift)) + skip
        out, st = pb.conv_gn([Act(hmid, H, W, rb.cout, st_h)], conv2.weight, conv2.bias, res, 1.0, True, gamma=n2.weight,
                             beta=n2.bias, groups=GN_GROUPS, eps=GN_EPS, silu=True, ada=self.ada, ada_stride=self.P,
                             ada_off=self.ada_off[id(rb)])
        return Act(out, H, W, rb.cout, st)

    def _attention(self, ab: _OAAttnP, x: Act) -> Act:
        """layout_unet_v1.py:416-532 (per-step part; the condition-only tensors come from set_condition)."""
        pb, plan, B = self.pb, self.plan, self.B
        C, nh, T = x.C, ab.num_heads, x.H * x.W
        L2 = 13
        wq = ab.qkv_projector.weight.detach().reshape(3 * C, C, 1, 1)
        qkv, _ = pb.conv_gn([x], wq, ab.qkv_projector.bias, None, 1.0, False, gamma=ab.norm_for_qkv.weight,
                            beta=ab.norm_for_qkv.bias, groups=GN_GROUPS, eps=GN_EPS, silu=False)
        bufs = dict(pos_p=plan.f32(B, T, C), kl=plan.f32(B, L2, C), pos_l=plan.f32(B, L2, C), vl=plan.f32(B, L2, C))
        res_key = f"image_patch_bbox_embedding_for_resolution{self.m.image_size // (self.H // x.H)}"
        self.attn_consts.append((ab, res_key, bufs))
        att = plan.operand(x.H, x.W, C)
        d = C // nh
        ws = torch.empty(max(int(self.lib.flash_attention_workspace(B, nh, T, L2, 2 * d, d)), 16), dtype=torch.uint8, device=self.dev)
        plan.bufs.append(ws)                # packed Q / K / V^T tile images of the tcgen05 attention
        plan.add(self.lib.flash_attention_oa, _ptr(qkv), _ptr(bufs["pos_p"]), _ptr(bufs["kl"]), _ptr(bufs["pos_l"]),
                 _ptr(bufs["vl"]), _ptr(att), x.W, plan.parts, B, C, nh, T, L2, 1.0 / math.sqrt(2 * d), _ptr(ws), name="attention_oa",
                 flops=2.0 * B * nh * T * (T + L2) * (3 * d))
        plan.flops += 2.0 * B * nh * T * (T + L2) * (3 * d)
        wo = ab.proj_out.weight.detach().reshape(C, C, 1, 1)
        out, st = pb.conv(att, x.H, x.W, wo, ab.proj_out.bias, x.t, 1.0, True)
        return Act(out, x.H, x.W, C, st)

    # ------------------------------------------------------------------------------------------
    @torch.no_grad()
    def set_condition(self, cond: dict) -> None:
        """Fold everything that depends only on the layout condition.  ALWAYS re-folds: callers decide the reuse --
        sample() / inpaint() fold once per call and then replay the step graph; forward() / p_step() fold on every call.
        (No address- or version-based memoisation: inference tensors carry no version counter and the caching allocator
        hands the same blocks to consecutive, different layouts.)  PyTorch ops on a few [B,64,13] / [B,64,T] tensors + one
        direct conv -- not on the per-step path."""
        m, B, dev = self.m, self.B, self.dev
        f = lambda t: t.to(dev, torch.float32)
        self.xf_proj.copy_(f(cond["xf_proj"]))
        if self.n_cond:
            cc = f(cond["concat_cond"])
            assert cc.shape[1] == self.n_cond, f"concat_cond has {cc.shape[1]} channels, model expects {self.n_cond}"
            self.cst_in[:, :, :self.n_cond] = cc.permute(0, 2, 3, 1).reshape(B, self.H * self.W, self.n_cond)
        self.lib.conv_direct_f32(_ptr(self.cst_in), _ptr(self.w_cst), _ptr(self.b0), _ptr(self.cst), B, self.H, self.W,
                                 self.cst_in.shape[-1], self.cst.shape[-1], 3, 1, _lib.current_stream(dev))
        xf_out, cls_e, box_e = f(cond["xf_out"]), f(cond["obj_class_embedding"]), f(cond["obj_bbox_embedding"])
        for ab, res_key, bufs in self.attn_consts:
            C = ab.channels
            pw, pbias = ab.layout_position_embedding_projector.weight.float(), ab.layout_position_embedding_projector.bias.float()
            gnp, gnl = ab.norm_for_image_patch_positional_embedding, ab.norm_for_layout_positional_embedding
            pos_p = F.group_norm(F.conv1d(f(cond[res_key]), pw, pbias), GN_GROUPS, gnp.weight.float(), gnp.bias.float(), GN_EPS)
            pos_l = F.group_norm(F.conv1d(box_e, pw, pbias), GN_GROUPS, gnl.weight.float(), gnl.bias.float(), GN_EPS)
            gnc = ab.norm_for_obj_class_embedding
            content = (xf_out + F.group_norm(cls_e, GN_GROUPS, gnc.weight.float(), gnc.bias.float(), GN_EPS)) / 2
            kv = F.conv1d(content, ab.layout_content_embedding_projector.weight.float(),
                          ab.layout_content_embedding_projector.bias.float())
            bufs["pos_p"].copy_(pos_p.transpose(1, 2))
            bufs["pos_l"].copy_(pos_l.transpose(1, 2))
            bufs["kl"].copy_(kv[:, :C].transpose(1, 2))
            bufs["vl"].copy_(kv[:, C:].transpose(1, 2))

    def launch(self, stream: int | None = None):
        if stream is None:
            stream = _lib.current_stream(self.dev)
        self.plan.run(stream)

    def __call__(self, x: torch.Tensor, log_snr: torch.Tensor) -> torch.Tensor:
        self.x_in.copy_(x)
        self.t_in.copy_(log_snr)
        self.launch()
        return self.pred.clone()
