"""CPU emulator of the libb200lidar C-ABI -- TEST INFRASTRUCTURE.

Each method takes exactly the arguments of the matching ``b200_*`` entry point (raw addresses of CPU
buffers instead of device pointers) and re-implements the documented contract with torch on the CPU.
It lets ``-m "not gpu"`` tests execute the HOST logic (plan construction, weight packing order, buffer
wiring, sampler loop) end to end against the oracle without a GPU.  It is never reachable from the
product path (only ``tests/`` call ``_lib.set_test_lib``).
"""
from __future__ import annotations

import ctypes
import math

import numpy as np
import torch
import torch.nn.functional as F

KC = 32


def _arr(ptr, n, ctype, npdtype):
    if ptr in (0, None):
        return None
    buf = (ctype * int(n)).from_address(int(ptr))
    return torch.from_numpy(np.frombuffer(buf, dtype=npdtype, count=int(n)))


def f32(ptr, *shape):
    t = _arr(ptr, math.prod(shape), ctypes.c_float, np.float32)
    return None if t is None else t.view(*shape)


def f16(ptr, *shape):
    t = _arr(ptr, math.prod(shape), ctypes.c_uint16, np.float16)
    return None if t is None else t.view(*shape)


def f64(ptr, *shape):
    t = _arr(ptr, math.prod(shape), ctypes.c_double, np.float64)
    return None if t is None else t.view(*shape)


def i32(ptr, *shape):
    t = _arr(ptr, math.prod(shape), ctypes.c_int32, np.int32)
    return None if t is None else t.view(*shape)


def u8(ptr, *shape):
    t = _arr(ptr, math.prod(shape), ctypes.c_uint8, np.uint8)
    return None if t is None else t.view(*shape)


F8_LO_SCALE = 2048.0


def e4m3(v):
    """fp32 -> e4m3 (round to nearest even, saturate to +-448) as uint8 codes"""
    return v.float().clamp(-448.0, 448.0).to(torch.float8_e4m3fn).view(torch.uint8)


def e4m3_value(codes):
    return codes.view(torch.float8_e4m3fn).float()


def n_planes(parts):
    return 1 if parts == 1 else 2


OPX, OTW = 130, 128     # pixels per operand slab (128 + 2 ring-halo pixels), tile width


def to_tiles(x, g=8):
    """[..., H, W, C] -> conv operand layout [..., H, W/128, C/g, 130, g]: per 128-pixel tile and g-channel group 130
    pixels = ring neighbour (w0-1) mod W, the 128 tile pixels, ring neighbour (w0+128) mod W  (csrc/common.cuh)"""
    *lead, H, W, C = x.shape
    WT = W // OTW

    def tiles(t):
        return t.reshape(*lead, H, WT, OTW, C // g, g)
    left = tiles(torch.roll(x, 1, dims=-2))[..., 0:1, :, :]
    right = tiles(torch.roll(x, -1, dims=-2))[..., OTW - 1:OTW, :, :]
    return torch.cat([left, tiles(x), right], dim=-3).transpose(-3, -2).contiguous()


def from_tiles(t):
    """inverse of to_tiles (reads the tile bodies): [..., H, WT, C/g, 130, g] -> [..., H, W, C]"""
    *lead, H, WT, G, P, g = t.shape
    return t[..., 1:OTW + 1, :].transpose(-3, -2).reshape(*lead, H, WT * OTW, G * g)


def operand_elems(B, H, W, C):
    """fp16-sized elements of ONE operand plane"""
    return B * H * (W // OTW) * (C // 8) * OPX * 8


def store_operand(ptr, x, parts, B, H, W, C):
    """x [B, H, W, C] fp32 -> conv operand at ``ptr`` (include/b200lidar.h, "conv operand layout")"""
    x = x.reshape(B, H, W, C).float()
    WT = W // OTW
    hi = x.half()
    f16(ptr, B, H, WT, C // 8, OPX, 8).copy_(to_tiles(hi))
    plane = 2 * operand_elems(B, H, W, C)
    if parts == 2:
        f16(ptr + plane, B, H, WT, C // 8, OPX, 8).copy_(to_tiles((x - hi.float()).half()))
    elif parts == 3:
        l8 = to_tiles(e4m3((x - hi.float()) * F8_LO_SCALE), 16)       # [B, H, WT, C/16, 130, 16]
        a8 = to_tiles(e4m3(x), 16)
        u8(ptr + plane, B, H, WT, C // 16, 2, OPX, 16).copy_(torch.stack([l8, a8], dim=4))


def load_operand(ptr, parts, B, H, W, C):
    """-> list of fp32 [B, H, W, C] tensors: [hi] / [hi, lo] / [hi, L8, A8]  (tile bodies)"""
    WT = W // OTW
    plane = 2 * operand_elems(B, H, W, C)
    out = [from_tiles(f16(ptr, B, H, WT, C // 8, OPX, 8)).float()]
    if parts == 2:
        out.append(from_tiles(f16(ptr + plane, B, H, WT, C // 8, OPX, 8)).float())
    elif parts == 3:
        pair = e4m3_value(u8(ptr + plane, B, H, WT, C // 16, 2, OPX, 16))
        for sub in range(2):
            out.append(from_tiles(pair[:, :, :, :, sub]))
    return out


def _ring_pad(x, pad, ring):
    if pad == 0:
        return x
    x = F.pad(x, (pad, pad, 0, 0), mode="circular" if ring else "constant")
    return F.pad(x, (0, 0, pad, pad))


class EmulatedLib:
    def __init__(self):
        self.n_launches = 0
        self.calls = []

    def _rec(self, name):
        self.n_launches += 1
        self.calls.append(name)

    # ---- weights ----
    @staticmethod
    def _split(v, parts):
        hi = v.half()
        if parts == 1:
            return hi[None]
        return torch.stack([hi, (v - hi.float()).half()])

    @staticmethod
    def conv_merged(bn, rows, parts):
        return 1 if (parts == 2 and 2 * rows * bn <= 256) else 0

    @staticmethod
    def _wplanes(v, parts):
        """scaled fp32 weights -> the planes the conv multiplies with (fp32 values):
        parts 1: [fp16(v)]; 2: [hi, lo]; 3: [hi, e4m3(v 2^-11), e4m3(v - hi)]"""
        hi = v.half().float()
        if parts == 1:
            return [hi]
        if parts == 2:
            return [hi, (v - hi).half().float()]
        return [hi, e4m3_value(e4m3(v / F8_LO_SCALE)), e4m3_value(e4m3(v - hi))]

    def pack_conv_weight(self, w, out, Cout, Cin, taps, bn, rows, parts, wscale, stream):
        self._rec("pack_conv_weight")
        v = f32(w, Cout, Cin, taps) * wscale
        if rows == 0:
            # column walk (csrc/conv_col.cuh): per (16-channel chunk c, dx) two planes of [2 k-groups][192 = dy * 64 + cout][16 B]:
            # plane 0 fp16(w s) (k-group = 8 channels), plane 1 e4m3 {w s 2^-11 | w s - fp16(w s)} (16 channels each)
            assert parts == 3 and taps == 9 and bn == 64 and Cin == 64 and Cout == 64
            hi = v.half().view(64, 4, 2, 8, 3, 3).permute(1, 5, 2, 4, 0, 3).contiguous()        # [c][dx][kg][dy][n][8]
            w1 = e4m3(v / F8_LO_SCALE).view(64, 4, 16, 3, 3)
            w2 = e4m3(v - v.half().float()).view(64, 4, 16, 3, 3)
            p1 = torch.stack([w1, w2], dim=0).permute(2, 5, 0, 4, 1, 3).contiguous()            # [c][dx][2][dy][n][16]
            img = torch.cat([hi.view(torch.uint8).reshape(4, 3, -1), p1.reshape(4, 3, -1)], dim=-1)
            u8(out, Cout * Cin * taps * 4).copy_(img.reshape(-1))
            return 0
        if parts >= 3:
            # per (n-tile, 16-channel chunk, tap): plane 0 [2][bn][8] fp16, plane 1 [2 (hi8, lo8)][bn][16] e4m3
            hi = v.half().view(Cout // bn, bn, Cin // 16, 2, 8, taps).permute(0, 2, 5, 3, 1, 4).contiguous()
            w1 = e4m3(v / F8_LO_SCALE).view(Cout // bn, bn, Cin // 16, 16, taps)
            w2 = e4m3(v - v.half().float()).view(Cout // bn, bn, Cin // 16, 16, taps)
            p1 = torch.stack([w1, w2], dim=0).permute(1, 3, 5, 0, 2, 4).contiguous()   # [nt][c][tap][2][bn][16]
            img = torch.cat([hi.view(torch.uint8).reshape(Cout // bn, Cin // 16, taps, -1),
                             p1.reshape(Cout // bn, Cin // 16, taps, -1)], dim=-1)
            u8(out, Cout * Cin * taps * 4).copy_(img.reshape(-1))
            return 0
        merged = self.conv_merged(bn, rows, parts)
        kc = 16 if parts == 2 else 32
        W = self._split(v, parts)        # [parts, Cout, Cin, taps]
        # -> [Cout/bn][Cin/kc][taps][parts][kc/8][bn][8]   (merged mode, parts == 2 and bn == 64: [..][kc/8][parts][bn][8])
        t = W.view(parts, Cout // bn, bn, Cin // kc, kc // 8, 8, taps)
        t = (t.permute(1, 3, 6, 4, 0, 2, 5) if merged else t.permute(1, 3, 6, 0, 4, 2, 5)).contiguous()
        f16(out, Cout * Cin * taps * parts).copy_(t.reshape(-1))
        return 0

    def pack_conv_weight_plain(self, w, out, Cout, Cin, taps, parts, wscale, stream):
        self._rec("pack_conv_weight_plain")
        W = torch.stack(self._wplanes(f32(w, Cout, Cin, taps) * wscale, parts)).half()
        f16(out, parts, taps, Cout, Cin).copy_(W.permute(0, 3, 1, 2))
        return 0

    def _conv(self, a, Wp, bias, res, scale, w_inv, out, stats, B, H, Wd, Cin, Cout, taps, ring, parts):
        """Wp: list of [Cout, Cin, k, k] fp32 weight planes matching load_operand()'s activation planes"""
        k = 3 if taps == 9 else 1
        xs = [t.permute(0, 3, 1, 2) for t in load_operand(a, parts, B, H, Wd, Cin)]
        if parts == 2:      # (a_hi + a_lo) x (w_hi + w_lo)
            xs, Wp = [xs[0] + xs[1]], [Wp[0] + Wp[1]]
        y = sum(F.conv2d(_ring_pad(x, k // 2, ring), w) for x, w in zip(xs, Wp)) * w_inv
        y = y.permute(0, 2, 3, 1)
        if bias:
            y = y + f32(bias, Cout)
        if res:
            y = y + f32(res, B, H, Wd, Cout)
        y = (y * scale).float()
        f32(out, B, H, Wd, Cout).copy_(y)
        if stats:
            st = f64(stats, B, Cout, 2)
            st[:, :, 0] += y.double().sum(dim=(1, 2))
            st[:, :, 1] += (y.double() ** 2).sum(dim=(1, 2))

    def conv_tc(self, a, wpacked, bias, res, scale, w_inv, out, stats, B, H, W, Cin, Cout, taps, ring, bn, rows,
                parts, stream):
        self._rec("conv_tc")
        assert W % 128 == 0 and Cin % KC == 0 and Cout % bn == 0 and H % rows == 0 and rows * bn <= 256
        k = 3 if taps == 9 else 1
        if parts >= 3:
            img = u8(wpacked, Cout // bn, Cin // 16, taps, bn * 64)
            hi = img[..., :bn * 32].contiguous().view(torch.float16).view(Cout // bn, Cin // 16, taps, 2, bn, 8)
            Wp = [hi.float().permute(0, 4, 1, 3, 5, 2).reshape(Cout, Cin, k, k)]
            p1 = e4m3_value(img[..., bn * 32:].contiguous()).view(Cout // bn, Cin // 16, taps, 2, bn, 16)
            for sub in range(2):
                Wp.append(p1[:, :, :, sub].permute(0, 3, 1, 4, 2).reshape(Cout, Cin, k, k))
            self._conv(a, Wp, bias, res, scale, w_inv, out, stats, B, H, W, Cin, Cout, taps, ring, 3)
            return 0
        kc = 16 if parts == 2 else 32
        if self.conv_merged(bn, rows, parts):
            t = f16(wpacked, Cout // bn, Cin // kc, taps, kc // 8, parts, bn, 8).float().permute(4, 0, 1, 2, 3, 5, 6)
        else:
            t = f16(wpacked, Cout // bn, Cin // kc, taps, parts, kc // 8, bn, 8).float().permute(3, 0, 1, 2, 4, 5, 6)
        Wp = [tp.permute(0, 4, 1, 3, 5, 2).reshape(Cout, Cin, k, k) for tp in t]
        self._conv(a, Wp, bias, res, scale, w_inv, out, stats, B, H, W, Cin, Cout, taps, ring, parts)
        return 0

    def conv_tc_splitk(self, a, wpacked, bias, res, scale, w_inv, out, stats, workspace, splits, B, H, W, Cin, Cout, taps,
                       ring, bn, rows, parts, stream):
        """split-K variant of conv_tc: same contract (the K slices are an implementation detail of the GPU kernel)"""
        assert splits in (2, 4, 8) and (Cin // (32 if parts == 1 else 16)) % splits == 0 and workspace
        self.conv_tc(a, wpacked, bias, res, scale, w_inv, out, stats, B, H, W, Cin, Cout, taps, ring, bn, rows, parts, stream)
        self.calls[-1] = "conv_tc_splitk"
        return 0

    def conv_gn_tc(self, x0, C0, x1, C1, st0, st1, gamma, beta, ada, ada_stride, groups, eps, silu, wpacked, bias, res,
                   scale, w_inv, out, stats, B, H, W, Cout, taps, ring, bn, rows, parts, stream):
        """GroupNorm(+AdaGN)-apply + SiLU + operand split fused in front of conv_tc: == gn_act_f16 into a scratch operand,
        then conv_tc (the kernel builds the same operand in shared memory)"""
        assert parts in (2, 3) and C0 % 16 == 0 and C1 % 16 == 0 and C0 + C1 <= 1024 and groups <= 32
        Cin = C0 + C1
        n0 = self.n_launches
        scratch = torch.zeros(n_planes(parts) * operand_elems(B, H, W, Cin), dtype=torch.float16)
        self.gn_act_f16(x0, C0, x1, C1, st0, st1, gamma, beta, ada, ada_stride, groups, eps, silu, scratch.data_ptr(), 0,
                        parts, B, H, W, stream)
        self.calls.pop()
        if rows == 0:      # column walk (csrc/conv_col.cuh): same contract, restricted shapes, its own weight image
            assert parts == 3 and taps == 9 and bn == 64 and C0 == 64 and C1 == 0 and Cout == 64
            img = u8(wpacked, 4, 3, 2, 2 * 192 * 16)
            hi = img[:, :, 0].contiguous().view(torch.float16).view(4, 3, 2, 3, 64, 8)          # [c][dx][kg][dy][n][8]
            Wp = [hi.float().permute(4, 0, 2, 5, 3, 1).reshape(64, 64, 3, 3)]
            p1 = e4m3_value(img[:, :, 1].contiguous()).view(4, 3, 2, 3, 64, 16)                 # [c][dx][sub][dy][n][16]
            for sub in range(2):
                Wp.append(p1[:, :, sub].permute(3, 0, 4, 2, 1).reshape(64, 64, 3, 3))
            self._rec("conv_tc")
            self._conv(scratch.data_ptr(), Wp, bias, res, scale, w_inv, out, stats, B, H, W, Cin, Cout, taps, ring, 3)
            self.calls[-1] = "conv_gn_tc"
            self.n_launches = n0 + 1
            return 0
        self.conv_tc(scratch.data_ptr(), wpacked, bias, res, scale, w_inv, out, stats, B, H, W, Cin, Cout, taps, ring, bn,
                     rows, parts, stream)
        self.calls[-1] = "conv_gn_tc"
        self.n_launches = n0 + 1
        return 0

    def conv_ffma(self, a, w16, bias, res, scale, w_inv, out, stats, B, H, W, Cin, Cout, taps, ring, parts, stream):
        self._rec("conv_ffma")
        k = 3 if taps == 9 else 1
        Wp = [t.permute(1, 2, 0).reshape(Cout, Cin, k, k) for t in f16(w16, parts, taps, Cout, Cin).float()]
        self._conv(a, Wp, bias, res, scale, w_inv, out, stats, B, H, W, Cin, Cout, taps, ring, parts)
        return 0

    # ---- GN / stats ----
    def gn_act_f16(self, x0, C0, x1, C1, st0, st1, gamma, beta, ada, ada_stride, groups, eps, silu, y, y_raw, parts, B,
                   H, W, stream):
        self._rec("gn_act_f16")
        HW = H * W
        x = f32(x0, B, HW, C0)
        if C1:
            x = torch.cat([x, f32(x1, B, HW, C1)], dim=-1)
        C = C0 + C1
        if y_raw:
            store_operand(y_raw, x, parts, B, H, W, C)
        if st0:
            st = f64(st0, B, C0, 2)
            if C1:
                st = torch.cat([st, f64(st1, B, C1, 2)], dim=1)
            cpg = C // groups
            g = st.view(B, groups, cpg, 2).sum(dim=2)
            n = HW * cpg
            mean = g[..., 0] / n
            var = (g[..., 1] / n - mean * mean).clamp(min=0)
            rstd = 1.0 / torch.sqrt(var + eps)
            mean = mean.float().repeat_interleave(cpg, dim=1)
            rstd = rstd.float().repeat_interleave(cpg, dim=1)
            a = rstd
            b = -mean * rstd
            if gamma:
                a = a * f32(gamma, C)
                b = b * f32(gamma, C) + f32(beta, C)
            if ada:
                full = _arr(ada, (B - 1) * ada_stride + 2 * C, ctypes.c_float, np.float32)
                rows = torch.stack([full[i * ada_stride:i * ada_stride + 2 * C] for i in range(B)])
                sc, sh = 1 + rows[:, :C], rows[:, C:]
                a = a * sc
                b = b * sc + sh
            x = x * a[:, None, :] + b[:, None, :]
        if silu:
            x = F.silu(x)
        store_operand(y, x, parts, B, H, W, C)
        return 0

    def gn_act_f32(self, x, st, gamma, beta, groups, eps, silu, y, B, HW, C, stream):
        self._rec("gn_act_f32")
        xt = f32(x, B, HW, C)
        s = f64(st, B, C, 2)
        cpg = C // groups
        g = s.view(B, groups, cpg, 2).sum(dim=2)
        n = HW * cpg
        mean = g[..., 0] / n
        var = (g[..., 1] / n - mean * mean).clamp(min=0)
        rstd = (1.0 / torch.sqrt(var + eps)).float().repeat_interleave(cpg, dim=1)
        mean = mean.float().repeat_interleave(cpg, dim=1)
        a, b = rstd, -mean * rstd
        if gamma:
            a = a * f32(gamma, C)
            b = b * f32(gamma, C) + f32(beta, C)
        o = xt * a[:, None, :] + b[:, None, :]
        if silu:
            o = F.silu(o)
        f32(y, B, HW, C).copy_(o)
        return 0

    def attention_oa(self, qkv, pos_p, kl, pos_l, vl, out, out_w, parts, B, C, heads, T, L2, scale2, stream):
        self._rec("attention_oa")
        d = C // heads
        QKV = f32(qkv, B, T, 3 * C)
        def hd(t, n):
            return t.reshape(B, n, heads, d).transpose(1, 2)       # B, h, n, d
        q, k, v = hd(QKV[..., :C], T), hd(QKV[..., C:2 * C], T), hd(QKV[..., 2 * C:], T)
        pp, pl = hd(f32(pos_p, B, T, C), T), hd(f32(pos_l, B, L2, C), L2)
        klh, vlh = hd(f32(kl, B, L2, C), L2), hd(f32(vl, B, L2, C), L2)
        qm = torch.cat([q, pp], -1)
        km = torch.cat([torch.cat([k, pp], -1), torch.cat([klh, pl], -1)], 2)
        vm = torch.cat([v, vlh], 2)
        att = torch.softmax(qm @ km.transpose(-1, -2) * scale2, dim=-1)
        o = (att @ vm).transpose(1, 2).reshape(B, T, C)
        store_operand(out, o, parts, B, T // out_w, out_w, C)
        return 0

    def channel_stats(self, x, stats, B, HW, C, stream):
        self._rec("channel_stats")
        t = f32(x, B, HW, C).double()
        st = f64(stats, B, C, 2)
        st[:, :, 0] += t.sum(1)
        st[:, :, 1] += (t * t).sum(1)
        return 0

    def fir_resample(self, x, y, stats, B, H, W, C, up, ring, stream):
        self._rec("fir_resample")
        from oracle import unet_torch as O
        t = f32(x, B, H, W, C).permute(0, 3, 1, 2)
        r = (O.fir_up2(t, bool(ring)) if up else O.fir_down2(t, bool(ring))).permute(0, 2, 3, 1).contiguous()
        f32(y, *r.shape).copy_(r)
        if stats:
            st = f64(stats, B, C, 2)
            st[:, :, 0] += r.double().sum(dim=(1, 2))
            st[:, :, 1] += (r.double() ** 2).sum(dim=(1, 2))
        return 0

    def fir_up_operand(self, x, y, parts, B, H, W, C, ring, stream):
        self._rec("fir_up_operand")
        from oracle import unet_torch as O
        t = f32(x, B, H, W, C).permute(0, 3, 1, 2)
        r = O.fir_up2(t, bool(ring)).permute(0, 2, 3, 1).contiguous()
        store_operand(y, r, parts, B, 2 * H, 2 * W, C)
        return 0

    def time_embed(self, t, w1, b1, w2, b2, add, wp, bp, temb, ada, B, Cs, E, P, stream):
        self._rec("time_embed")
        tv = f32(t, B)
        half = Cs // 2
        fr = torch.exp(-math.log(10000.0) / (half - 1) * torch.arange(half, dtype=torch.float32))
        a = tv[:, None] * fr[None]
        e = torch.cat([a.sin(), a.cos()], -1)
        h = F.silu(F.linear(e, f32(w1, E, Cs), f32(b1, E)))
        te = F.linear(h, f32(w2, E, E), f32(b2, E))
        if add:
            te = te + f32(add, B, E)
        f32(temb, B, E).copy_(te)
        if P > 0:
            f32(ada, B, P).copy_(F.linear(F.silu(te), f32(wp, P, E), f32(bp, P)))
        return 0

    def in_conv(self, x, w, cst, cst_batched, out, stats, B, H, W, Cx, Cout, ring, stream):
        self._rec("in_conv")
        xt = f32(x, B, Cx, H, W)
        y = F.conv2d(_ring_pad(xt, 1, ring), f32(w, Cout, Cx, 3, 3)).permute(0, 2, 3, 1)
        y = (y + f32(cst, B if cst_batched else 1, H, W, Cout)).contiguous()
        f32(out, B, H, W, Cout).copy_(y)
        if stats:
            st = f64(stats, B, Cout, 2)
            st[:, :, 0] += y.double().sum(dim=(1, 2))
            st[:, :, 1] += (y.double() ** 2).sum(dim=(1, 2))
        return 0

    def conv_direct_f32(self, x, w, bias, out, B, H, W, Cin, Cout, k, ring, stream):
        self._rec("conv_direct_f32")
        xt = f32(x, B, H, W, Cin).permute(0, 3, 1, 2)
        y = F.conv2d(_ring_pad(xt, k // 2, ring), f32(w, Cout, Cin, k, k), f32(bias, Cout) if bias else None)
        f32(out, B, H, W, Cout).copy_(y.permute(0, 2, 3, 1))
        return 0

    def out_conv(self, a, a_is_f16, w, bias, pred, B, H, W, Cin, Cout, ring, stream):
        self._rec("out_conv")
        xt = (f16(a, B, H, W, Cin).float() if a_is_f16 else f32(a, B, H, W, Cin)).permute(0, 3, 1, 2)
        y = F.conv2d(_ring_pad(xt, 1, ring), f32(w, Cout, Cin, 3, 3), f32(bias, Cout))
        f32(pred, B, Cout, H, W).copy_(y)
        return 0

    def attention(self, q, ldq, qoff, k, ldk, koff, v, ldv, voff, out, ldo, out_w, parts, B, heads, Tq, Tk, dqk, dv,
                  scale, stream):
        self._rec("attention")
        Q = f32(q, B, Tq, ldq)[:, :, qoff:qoff + heads * dqk].reshape(B, Tq, heads, dqk).transpose(1, 2)
        K = f32(k, B, Tk, ldk)[:, :, koff:koff + heads * dqk].reshape(B, Tk, heads, dqk).transpose(1, 2)
        V = f32(v, B, Tk, ldv)[:, :, voff:voff + heads * dv].reshape(B, Tk, heads, dv).transpose(1, 2)
        att = torch.softmax(Q @ K.transpose(-1, -2) * scale, dim=-1)
        o = (att @ V).transpose(1, 2).reshape(B, Tq, heads * dv)
        assert ldo == heads * dv
        store_operand(out, o, parts, B, Tq // out_w, out_w, ldo)
        return 0

    @staticmethod
    def flash_attention_workspace(B, heads, T, Tx, dq, dv):
        return 16

    def flash_attention(self, qkv, E, out, out_w, parts, B, heads, T, scale, workspace, stream):
        d = E // heads
        return self.attention(qkv, 3 * E, 0, qkv, 3 * E, E, qkv, 3 * E, 2 * E, out, E, out_w, parts, B, heads, T, T, d, d,
                              scale, stream)

    def flash_attention_oa(self, *a):
        return self.attention_oa(*a[:-2], a[-1])            # (workspace is scratch of the device kernel)

    def sampler_update(self, x_t, pred, noise, coef, x_s, B, n, mode, objective, clip, stream):
        self._rec("sampler_update")
        xt, pr = f32(x_t, B, n).clone(), f32(pred, B, n)
        c = f32(coef, B, 8)
        a_t, s_t, a_s, s_s, c1, c2, cc = [c[:, i:i + 1] for i in range(7)]
        x0 = (xt - s_t * pr) / a_t if objective == 0 else (a_t * xt - s_t * pr if objective == 1 else pr)
        if clip > 0:
            x0 = x0.clamp(-clip, clip)
        if mode == 0:
            eps = (xt - a_t * x0) / s_t
            o = a_s * x0 + c2 * eps
            if noise:
                o = o + c1 * f32(noise, B, n)
        else:
            o = a_s * (xt * (1 - cc) / a_t + cc * x0) + s_s * cc.sqrt() * f32(noise, B, n)
        f32(x_s, B, n).copy_(o)
        return 0


# ---------------------------------------------------------------------------------------------------------
# point-cloud side of the ABI (K6 / K7 / K8): backed by the C oracle (oracle/lidar_ops.c, oracle/metrics_ops.c),
# so that the HOST mirrors (lidarcrafter_b200/ops.py, metric_utils.py: batching, bounds, list / dtype handling,
# workspace sizing) run end to end without a GPU.
# ---------------------------------------------------------------------------------------------------------
def i64(ptr, *shape):
    t = _arr(ptr, math.prod(shape), ctypes.c_int64, np.int64)
    return None if t is None else t.view(*shape)


def _host_ints(obj, n):
    """a ctypes int32 array (host pointer argument of the ABI) -> list of python ints"""
    return [int(obj[i]) for i in range(n)]


def _install_pointcloud_entries():
    from oracle import lidar_ops as LO
    from oracle import metrics_ops as MO

    def range_project(self, points, npts, out, grid, zbuf, F, M, H, W, min_d, max_d, fov_up, fov_down, stream):
        self._rec("range_project")
        pts, o, n = f32(points, F, M, 4), f32(out, F, H, W, 6), i32(npts, F)
        g = i32(grid, F, M, 2)
        for f in range(F):
            cnt = M if n is None else int(n[f])
            img, gr, _ = LO.range_project(pts[f, :cnt].numpy(), H, W, min_d, max_d, fov_up, fov_down)
            o[f] = torch.from_numpy(img)
            if g is not None:
                g[f, :cnt] = torch.from_numpy(gr)
        return 0

    def range_project_f64(self, points, npts, out, zbuf, winner, F, M, H, W, min_d, max_d, fov_up, fov_down, stream):
        self._rec("range_project_f64")
        pts, o, n = f64(points, F, M, 4), f32(out, F, H, W, 6), i32(npts, F)
        for f in range(F):
            cnt = M if n is None else int(n[f])
            img, _ = LO.range_project_f64(pts[f, :cnt].numpy(), H, W, min_d, max_d, fov_up, fov_down)
            o[f] = torch.from_numpy(img)
        return 0

    def boxes_to_mask_workspace(self, F, N):
        return 32 * F * N

    def boxes_to_mask(self, boxes, is64, F, N, H, W, fov_up, fov_down, boxes_2d, mask, weight, ws, stream):
        self._rec("boxes_to_mask")
        bx = (f64 if is64 else f32)(boxes, F, N, 8)
        b2, m, w = f64(boxes_2d, F, N, 4), f32(mask, F, 2, H, W), f32(weight, F, H, W)
        for f in range(F):
            r2, rm, rw = LO.boxes_to_mask(bx[f].numpy(), H, W, fov_up, fov_down)
            b2[f], m[f] = torch.from_numpy(r2), torch.from_numpy(rm)
            if w is not None:
                w[f] = torch.from_numpy(rw)
        return 0

    def points_in_boxes(self, pts, boxes, out, N, M, stream):
        self._rec("points_in_boxes")
        i32(out, N, M).copy_(torch.from_numpy(LO.points_in_boxes(f32(pts, M, 3).numpy(), f32(boxes, N, 7).numpy())))
        return 0

    def points_in_boxes_first(self, pts, boxes, out, B, N, M, stream):
        self._rec("points_in_boxes_first")
        i32(out, B, M).copy_(torch.from_numpy(LO.points_in_boxes_first(f32(pts, B, M, 3).numpy(), f32(boxes, B, N, 7).numpy())))
        return 0

    def voxel_index(self, pts, rois, out, N, M, ox, oy, oz, stream):
        self._rec("voxel_index")
        i32(out, N, M).copy_(torch.from_numpy(LO.voxel_index(f32(pts, M, 3).numpy(), f32(rois, N, 7).numpy(), (ox, oy, oz))))
        return 0

    def pcd2range(self, pcd, npts, feature, proj_range, proj_feature, zbuf, F, M, H, W, fov_up, fov_down, dmin, dmax, fill,
                  stream):
        self._rec("pcd2range")
        p, ft = f32(pcd, F, M, 3), f32(feature, F, M)
        r, pf = f32(proj_range, F, H, W), f32(proj_feature, F, H, W)
        for f in range(F):
            rr, ff, _ = MO.pcd2range(p[f].numpy(), (H, W), (fov_up, fov_down), (dmin, dmax),
                                     feature=None if ft is None else ft[f].numpy(), fill=fill)
            r[f] = torch.from_numpy(rr)
            if pf is not None:
                pf[f] = torch.from_numpy(ff)
        return 0

    def range2xyz(self, img, xyz, F, H, W, fov_up, fov_down, dmin, dmax, depth_scale, log_scale, stream):
        self._rec("range2xyz")
        im, o = f32(img, F, H, W), f64(xyz, F, 3, H, W)
        for f in range(F):
            o[f] = torch.from_numpy(MO.range2xyz(im[f].numpy(), (fov_up, fov_down), (dmin, dmax), depth_scale, bool(log_scale)))
        return 0

    def quantize_coords(self, coords, is_f64, M, D, stride, v0, v1, v2, div_f32, voxel, minmax, stream):
        self._rec("quantize_coords")
        c = (f64 if is_f64 else f32)(coords, M, stride)[:, :D].numpy()
        v = MO.quantize(np.ascontiguousarray(c), (v0, v1, v2)[:D], bool(div_f32))
        i32(voxel, M, D).copy_(torch.from_numpy(v))
        mm = i32(minmax, 6)
        mm[:D] = torch.from_numpy(v.min(0))
        mm[3:3 + D] = torch.from_numpy(v.max(0))
        return 0

    def sparse_quantize_workspace(self, minmax_host, D, M):
        mm = _host_ints(minmax_host, 6)
        cells = 1
        for d in range(D):
            cells *= mm[3 + d] - mm[d] + 1
        return 0 if cells > (1 << 35) else 256 + 8 * M       # the emulator sorts; it only mirrors the size limit

    def ravel_hash(self, voxel, M, D, minmax_host, out, stream):
        self._rec("ravel_hash")
        keys, _, _, _ = MO.sparse_quantize(i32(voxel, M, D).numpy())
        i64(out, M).copy_(torch.from_numpy(keys.view(np.int64)))
        return 0

    def sparse_quantize(self, voxel, M, D, minmax_host, ws, ws_bytes, uniq, indices, inverse, n_unique, stream):
        self._rec("sparse_quantize")
        _, u, idx, inv = MO.sparse_quantize(i32(voxel, M, D).numpy())
        n = len(u)
        i32(n_unique, 1)[0] = n
        if uniq:
            i32(uniq, M, D)[:n] = torch.from_numpy(u)
        if indices:
            i64(indices, M)[:n] = torch.from_numpy(idx)
        if inverse:
            i64(inverse, M).copy_(torch.from_numpy(inv))
        return 0

    def bev_occupancy_sum(self, pcd, offsets, n_clouds, max_pts, stride, x0, x1, y0, y1, voxel, mbx, mby, X, Y, bitmap_ws,
                          volume_sum, stream):
        self._rec("bev_occupancy_sum")
        off = i32(offsets, n_clouds + 1).numpy()
        base = int(off[0])
        pts = f32(pcd, int(off[-1]), stride).numpy()
        clouds = [pts[off[c]:off[c + 1]] for c in range(n_clouds)]
        lo, hi = np.array([x0, y0], np.float32), np.array([x1, y1], np.float32)
        vol = np.zeros((X, Y), np.float32)
        cat = np.ascontiguousarray(np.concatenate(clouds), dtype=np.float32) if clouds else np.zeros((0, stride), np.float32)
        o = (off - base).astype(np.int32)
        LO.lib().oracle_bev_sum(LO._p(cat), LO._p(o), n_clouds, stride, LO._p(lo), LO._p(hi), ctypes.c_float(voxel),
                                LO._p(np.array([mbx, mby], np.int32)), LO._p(np.array([X, Y], np.int32)), LO._p(vol))
        f32(volume_sum, X, Y).add_(torch.from_numpy(vol))
        return 0

    def voxel_occupancy(self, pcd, M, stride, lohi_host, voxel, minb_host, dims_host, vol, stream):
        self._rec("voxel_occupancy")
        lohi = [float(lohi_host[i]) for i in range(6)]
        minb, dims = _host_ints(minb_host, 3), _host_ints(dims_host, 3)
        pts = np.ascontiguousarray(f32(pcd, M, stride).numpy(), dtype=np.float32)
        out = np.zeros(dims, np.float32)
        LO.lib().oracle_voxel_full(LO._p(pts), M, stride, LO._p(np.array(lohi[0::2], np.float32)),
                                   LO._p(np.array(lohi[1::2], np.float32)), ctypes.c_float(voxel),
                                   LO._p(np.array(minb, np.int32)), LO._p(np.array(dims, np.int32)), LO._p(out))
        f32(vol, *dims).copy_(torch.from_numpy(out))
        return 0

    for fn in (range_project, range_project_f64, boxes_to_mask_workspace, boxes_to_mask, points_in_boxes, points_in_boxes_first, voxel_index, pcd2range, range2xyz, quantize_coords,
               sparse_quantize_workspace, ravel_hash, sparse_quantize, bev_occupancy_sum, voxel_occupancy):
        setattr(EmulatedLib, fn.__name__, fn)


_install_pointcloud_entries()
