"""GPU parity tests, kernel by kernel, THROUGH THE C-ABI: every libb200lidar entry point is run on the
B200 with seeded inputs and compared with tests/abi_emulator.py (a CPU restatement of each kernel's
contract built on the same torch ops the oracle uses)."""
import math

import pytest
import torch

from abi_emulator import OPX, OTW, EmulatedLib, load_operand, n_planes, operand_elems, store_operand

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)


def _lib():
    from lidarcrafter_b200 import _lib as L
    assert L._TEST_LIB is None
    L.require_b200(0)
    return L.get_lib()


class Both:
    """Run one entry point on the GPU library and on the emulator; tensors are given as CPU tensors
    (inputs and zero-filled outputs) and mirrored to the device."""

    def __init__(self):
        self.gpu = _lib()
        self.emu = EmulatedLib()
        self.cpu, self.dev = [], []

    def t(self, x):
        x = x.contiguous()
        self.cpu.append(x)
        self.dev.append(x.cuda())
        return len(self.cpu) - 1

    def call(self, name, args):
        """args: ints/floats, or ('t', idx) tensor handles, or None."""
        ga = [self.dev[a[1]].data_ptr() if isinstance(a, tuple) else (0 if a is None else a) for a in args]
        ca = [self.cpu[a[1]].data_ptr() if isinstance(a, tuple) else (0 if a is None else a) for a in args]
        getattr(self.gpu, name)(*ga, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        getattr(self.emu, name)(*ca, 0)

    def out(self, idx):
        return self.dev[idx].cpu(), self.cpu[idx]


def rel(a, b):
    return float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))


def randn(*shape, seed=0, scale=1.0):
    return torch.randn(*shape, generator=torch.Generator().manual_seed(seed)) * scale


def operand_zeros(parts, B, H, W, C):
    """empty conv operand: [planes][B][H][W/128][C/8][130][8] fp16-sized elements (plane 1 of parts 3: e4m3 pairs)"""
    return torch.zeros(n_planes(parts), operand_elems(B, H, W, C), dtype=torch.float16)


def split(v, parts):
    """fp32 [B,H,W,C] -> conv operand (include/b200lidar.h "conv operand layout")"""
    B, H, W, C = v.shape
    t = operand_zeros(parts, B, H, W, C)
    store_operand(t.data_ptr(), v, parts, B, H, W, C)
    return t


def operand_close(g, c, parts, B, H, W, C):
    """compare two conv operands by VALUE: hi (+ lo) reconstruct the fp32 input; for parts 3 additionally the A8 plane"""
    pg, pc = load_operand(g.data_ptr(), parts, B, H, W, C), load_operand(c.data_ptr(), parts, B, H, W, C)
    # halo pixels of every slab duplicate the ring neighbours' body pixels, bit for bit (all planes)
    raw = g.view(torch.int16).view(n_planes(parts), B, H, W // OTW, C // 8, OPX, 8)
    assert torch.equal(raw[..., 0, :], torch.roll(raw, 1, dims=3)[..., OTW, :])
    assert torch.equal(raw[..., OPX - 1, :], torch.roll(raw, -1, dims=3)[..., 1, :])
    if parts == 1:
        assert rel(pg[0], pc[0]) < 6e-4
    elif parts == 2:
        assert rel(pg[0] + pg[1], pc[0] + pc[1]) < 2e-5
    else:
        assert rel(pg[0] + pg[1] / 2048.0, pc[0] + pc[1] / 2048.0) < 6e-5     # L8 keeps 3-4 bits of the residual
        assert rel(pg[2], pc[2]) < 2e-2                                         # A8 = e4m3(x): a few 1-ulp flips
        assert rel(pg[2], pc[0]) < 5e-2                                         # ... and it does encode x


def wp_zeros(Cout, Cin, taps, parts):
    return torch.zeros(Cout * Cin * taps * n_planes(parts), dtype=torch.float16)


# ------------------------------------------------------------------------------------------------
@pytest.mark.parametrize("parts", [2, 1, 3])
@pytest.mark.parametrize("taps", [9, 1])
@pytest.mark.parametrize("bn,rows,Cin,Cout,H,W,B", [
    (64, 2, 64, 64, 8, 256, 2),
    (64, 4, 64, 64, 8, 256, 2),
    (64, 2, 64, 64, 32, 1024, 3),        # > 148 tiles: persistent loop + TMEM double buffering
    (64, 2, 96, 64, 4, 128, 1),
    (64, 1, 32, 128, 3, 128, 2),
    (128, 2, 128, 128, 4, 256, 1),
    (128, 2, 64, 256, 16, 512, 2),
    (128, 1, 256, 128, 1, 128, 3),
])
def test_conv_tc_matches_contract(parts, taps, bn, rows, Cin, Cout, H, W, B):
    """tcgen05 implicit-GEMM conv (+bias, +residual, *scale, +stats) vs F.conv2d on the same fp16 operands."""
    h = Both()
    k = 3 if taps == 9 else 1
    w = h.t(randn(Cout, Cin, k, k, seed=1, scale=1 / math.sqrt(Cin * taps)))
    a = h.t(split(randn(B, H, W, Cin, seed=2), parts))
    bias = h.t(randn(Cout, seed=3, scale=0.1))
    res = h.t(randn(B, H, W, Cout, seed=4))
    wp = h.t(wp_zeros(Cout, Cin, taps, parts))
    out = h.t(torch.zeros(B, H, W, Cout))
    st = h.t(torch.zeros(B, Cout, 2, dtype=torch.float64))
    wscale = 64.0 if parts < 3 else 2.0 ** 16
    h.call("pack_conv_weight", [("t", w), ("t", wp), Cout, Cin, taps, bn, rows, parts, wscale])
    g, c = h.out(wp)
    assert torch.equal(g, c), "packed weight image differs"
    h.call("conv_tc", [("t", a), ("t", wp), ("t", bias), ("t", res), 0.70710678, 1.0 / wscale, ("t", out), ("t", st),
                       B, H, W, Cin, Cout, taps, 1, bn, rows, parts])
    g, c = h.out(out)
    assert rel(g, c) < 1e-5, rel(g, c)      # fp32 accumulation order differs (tensor core vs CPU conv)
    gs, cs = h.out(st)
    assert rel(gs, cs) < 2e-5


@pytest.mark.parametrize("parts", [2, 3])
@pytest.mark.parametrize("taps,bn,rows,C0,C1,Cout,H,W,B,groups,mode,ring", [
    (9, 64, 2, 64, 0, 64, 8, 256, 2, 8, "gn", 1),
    (9, 64, 2, 64, 0, 64, 32, 1024, 3, 8, "ada", 1),       # several tiles per CTA, sample changes inside a CTA's tile range
    (9, 64, 4, 64, 64, 64, 8, 256, 2, 8, "gn", 1),         # channel concat (up path), 4-row tiles
    (9, 64, 1, 32, 32, 64, 4, 128, 1, 8, "gn", 0),         # zero padding in W, concat inside one GroupNorm group layout
    (9, 128, 2, 128, 0, 256, 16, 512, 2, 32, "gn_ada", 1),
    (9, 128, 1, 256, 256, 128, 4, 128, 3, 8, "gn", 1),
    (9, 128, 1, 512, 512, 512, 4, 128, 1, 32, "gn_ada", 1),    # Cin = 1024 (LayoutUnetV1 output block 0)
    (1, 64, 4, 128, 0, 64, 8, 256, 2, 8, "raw", 1),        # 1x1 skip conv on the un-normalised input
    (1, 128, 2, 256, 0, 768, 4, 128, 2, 8, "gn_nosilu", 1),    # GroupNorm -> QKV projection
    (1, 64, 1, 64, 64, 64, 4, 128, 2, 8, "raw", 1),
    (1, 64, 2, 64, 0, 128, 4, 256, 1, 8, "raw", 1),
    (1, 128, 1, 512, 0, 128, 4, 128, 2, 32, "gn_nosilu", 1),
    # rows = 0: column walk (conv_col.cuh), 64 -> 64 channels, fp16f8 only
    (9, 64, 0, 64, 0, 64, 8, 256, 2, 8, "gn", 1),
    (9, 64, 0, 64, 0, 64, 32, 1024, 3, 8, "ada", 1),       # ~5 rows per CTA, column and sample changes inside a CTA's run
    (9, 64, 0, 64, 0, 64, 32, 1024, 8, 8, "gn_ada", 1),    # the bench shape
    (9, 64, 0, 64, 0, 64, 5, 128, 1, 8, "gn", 0),          # zero padding in W, odd H, one row per CTA
    (9, 64, 0, 64, 0, 64, 16, 512, 2, 32, "raw", 1),
    (9, 64, 0, 64, 0, 64, 64, 128, 4, 8, "gn_nosilu", 1),  # H = 64: long columns
])
def test_conv_gn_tc_fused_front(parts, taps, bn, rows, C0, C1, Cout, H, W, B, groups, mode, ring):
    """GroupNorm(+AdaGN)-apply + SiLU + operand split by the conv kernel's transform warps (b200_conv_gn_tc) vs the
    emulator (gn_act_f16 -> conv_tc on the CPU), and vs the separate GPU launches gn_act_f16 -> conv_tc: the fused kernel
    builds the same operand bits in shared memory and issues the same MMAs, so the outputs must be bit-identical."""
    if rows == 0 and parts != 3:
        pytest.skip("column walk: fp16f8 operands only")
    h = Both()
    Cin = C0 + C1
    k = 3 if taps == 9 else 1
    w = h.t(randn(Cout, Cin, k, k, seed=1, scale=1 / math.sqrt(Cin * taps)))
    x0 = h.t(randn(B, H, W, C0, seed=2) * 1.7 + 0.3)
    x1 = h.t(randn(B, H, W, C1, seed=8) * 0.6 - 0.2) if C1 else None
    bias = h.t(randn(Cout, seed=3, scale=0.1))
    res = h.t(randn(B, H, W, Cout, seed=4))
    wp = h.t(wp_zeros(Cout, Cin, taps, parts))
    out = h.t(torch.zeros(B, H, W, Cout))
    out2 = h.t(torch.zeros(B, H, W, Cout))
    st = h.t(torch.zeros(B, Cout, 2, dtype=torch.float64))
    st2 = h.t(torch.zeros(B, Cout, 2, dtype=torch.float64))
    s0 = h.t(torch.zeros(B, C0, 2, dtype=torch.float64))
    s1 = h.t(torch.zeros(B, C1, 2, dtype=torch.float64)) if C1 else None
    gam, bet = h.t(1 + 0.1 * randn(Cin, seed=5)), h.t(0.1 * randn(Cin, seed=6))
    P = 2 * Cin + 16
    ada = h.t(0.3 * randn(B, P, seed=7))
    y = h.t(operand_zeros(parts, B, H, W, Cin))
    wscale = 64.0 if parts < 3 else 2.0 ** 16
    norm = mode != "raw"
    affine = mode in ("gn", "gn_ada", "gn_nosilu")
    use_ada = mode in ("ada", "gn_ada")
    silu = 0 if mode in ("raw", "gn_nosilu") else 1
    h.call("pack_conv_weight", [("t", w), ("t", wp), Cout, Cin, taps, bn, rows, parts, wscale])
    if norm:
        h.call("channel_stats", [("t", x0), ("t", s0), B, H * W, C0])
        if C1:
            h.call("channel_stats", [("t", x1), ("t", s1), B, H * W, C1])
    front = [("t", x0), C0, ("t", x1) if C1 else None, C1, ("t", s0) if norm else None,
             ("t", s1) if (norm and C1) else None, ("t", gam) if affine else None, ("t", bet) if affine else None,
             ("t", ada) if use_ada else None, P, groups, 1e-6, silu]
    h.call("conv_gn_tc", front + [("t", wp), ("t", bias), ("t", res), 0.70710678, 1.0 / wscale, ("t", out), ("t", st), B, H, W,
                                  Cout, taps, ring, bn, rows, parts])
    g, c = h.out(out)
    assert rel(g, c) < 2e-5, rel(g, c)
    assert rel(*h.out(st)) < 2e-5
    # the separate launches on the GPU
    rows2 = rows if rows else (2 if H % 2 == 0 else 1)
    if rows == 0:       # the column walk has its own weight image: repack for the tile walk
        h.call("pack_conv_weight", [("t", w), ("t", wp), Cout, Cin, taps, bn, rows2, parts, wscale])
    h.call("gn_act_f16", front + [("t", y), None, parts, B, H, W])
    h.call("conv_tc", [("t", y), ("t", wp), ("t", bias), ("t", res), 0.70710678, 1.0 / wscale, ("t", out2), ("t", st2), B, H, W,
                       Cin, Cout, taps, ring, bn, rows2, parts])
    g2, _ = h.out(out2)
    if rows == 0:     # same operand bits and MMAs, but issued filter row by filter row: another fp32 accumulation order
        assert rel(g, g2) < 2e-6, rel(g, g2)
        assert rel(h.out(st)[0], h.out(st2)[0]) < 1e-6
    else:
        assert torch.equal(g, g2), float((g - g2).abs().max())


@pytest.mark.parametrize("parts,taps,bn,rows,Cin,Cout,H,W,B,splits", [
    (3, 9, 64, 1, 512, 512, 4, 128, 1, 4),        # the B = 1 deep level: 32 tiles, K = 4608
    (3, 9, 128, 1, 256, 256, 8, 256, 1, 2),
    (2, 9, 64, 1, 256, 256, 4, 128, 2, 8),
    (3, 1, 64, 1, 512, 256, 4, 128, 1, 4),
    (1, 9, 128, 2, 256, 128, 4, 128, 3, 2),
    (3, 9, 64, 2, 128, 64, 8, 256, 1, 8),
])
def test_conv_tc_splitk_matches_conv_tc(parts, taps, bn, rows, Cin, Cout, H, W, B, splits):
    """b200_conv_tc_splitk (K slices on `splits` CTAs per tile + reduce / epilogue kernel) vs the emulator and vs b200_conv_tc on
    the GPU: same operands, same epilogue, another fp32 summation order"""
    h = Both()
    k = 3 if taps == 9 else 1
    w = h.t(randn(Cout, Cin, k, k, seed=1, scale=1 / math.sqrt(Cin * taps)))
    a = h.t(split(randn(B, H, W, Cin, seed=2), parts))
    bias = h.t(randn(Cout, seed=3, scale=0.1))
    res = h.t(randn(B, H, W, Cout, seed=4))
    wp = h.t(wp_zeros(Cout, Cin, taps, parts))
    out = h.t(torch.zeros(B, H, W, Cout))
    out2 = h.t(torch.zeros(B, H, W, Cout))
    st = h.t(torch.zeros(B, Cout, 2, dtype=torch.float64))
    st2 = h.t(torch.zeros(B, Cout, 2, dtype=torch.float64))
    ws = h.t(torch.zeros(splits, B * H * W * Cout))
    wscale = 64.0 if parts < 3 else 2.0 ** 16
    h.call("pack_conv_weight", [("t", w), ("t", wp), Cout, Cin, taps, bn, rows, parts, wscale])
    h.call("conv_tc_splitk", [("t", a), ("t", wp), ("t", bias), ("t", res), 0.70710678, 1.0 / wscale, ("t", out), ("t", st),
                              ("t", ws), splits, B, H, W, Cin, Cout, taps, 1, bn, rows, parts])
    g, c = h.out(out)
    assert rel(g, c) < 1e-5, rel(g, c)
    assert rel(*h.out(st)) < 2e-5
    h.call("conv_tc", [("t", a), ("t", wp), ("t", bias), ("t", res), 0.70710678, 1.0 / wscale, ("t", out2), ("t", st2),
                       B, H, W, Cin, Cout, taps, 1, bn, rows, parts])
    g2, _ = h.out(out2)
    assert rel(g, g2) < 2e-5, rel(g, g2)       # K up to 4608 products summed in another order (fp32: ~sqrt(K) 2^-24)
    assert rel(h.out(st)[0], h.out(st2)[0]) < 1e-5
    # no bias / residual / statistics
    h.call("conv_tc_splitk", [("t", a), ("t", wp), None, None, 1.0, 1.0 / wscale, ("t", out), None, ("t", ws), splits, B, H, W,
                              Cin, Cout, taps, 0, bn, rows, parts])
    g, c = h.out(out)
    assert rel(g, c) < 1e-5, rel(g, c)


@pytest.mark.parametrize("H,W,B,with_stats", [(1, 128, 1, False), (2, 256, 3, True), (32, 1024, 1, True), (16, 512, 16, False)])
def test_conv_col_optional_arguments_and_borders(H, W, B, with_stats):
    """column walk (b200_conv_gn_tc rows = 0) without bias / residual / statistics / normalisation, single-row images, one CTA
    per row tile, and runs that span two samples: vs the emulator"""
    h = Both()
    w = h.t(randn(64, 64, 3, 3, seed=1, scale=1 / math.sqrt(64 * 9)))
    x0 = h.t(randn(B, H, W, 64, seed=2) * 1.3 - 0.1)
    wp = h.t(wp_zeros(64, 64, 9, 3))
    out = h.t(torch.zeros(B, H, W, 64))
    st = h.t(torch.zeros(B, 64, 2, dtype=torch.float64))
    wscale = 2.0 ** 16
    h.call("pack_conv_weight", [("t", w), ("t", wp), 64, 64, 9, 64, 0, 3, wscale])
    g, c = h.out(wp)
    assert torch.equal(g, c), "packed weight image differs"
    h.call("conv_gn_tc", [("t", x0), 64, None, 0, None, None, None, None, None, 0, 8, 1e-6, 1, ("t", wp), None, None, 1.0,
                          1.0 / wscale, ("t", out), ("t", st) if with_stats else None, B, H, W, 64, 9, 1, 64, 0, 3])
    g, c = h.out(out)
    assert rel(g, c) < 2e-5, rel(g, c)
    if with_stats:
        assert rel(*h.out(st)) < 2e-5


def test_conv_tc_zero_pad_no_bias_no_res():
    h = Both()
    B, H, W, Cin, Cout = 1, 4, 128, 64, 64
    w = h.t(randn(Cout, Cin, 3, 3, seed=1, scale=0.05))
    a = h.t(split(randn(B, H, W, Cin, seed=2), 2))
    wp = h.t(torch.zeros(Cout * Cin * 9 * 2, dtype=torch.float16))
    out = h.t(torch.zeros(B, H, W, Cout))
    h.call("pack_conv_weight", [("t", w), ("t", wp), Cout, Cin, 9, 64, 2, 2, 128.0])
    h.call("conv_tc", [("t", a), ("t", wp), None, None, 1.0, 1.0 / 128.0, ("t", out), None, B, H, W, Cin, Cout, 9, 0, 64,
                       2, 2])  # ring = 0: zero padding in W as well
    g, c = h.out(out)
    assert rel(g, c) < 2e-6


def test_conv_tc_split_reaches_fp32_accuracy():
    """fp16x3 split vs an fp64 convolution of the ORIGINAL fp32 operands: ~1e-6, single fp16 ~5e-4."""
    import torch.nn.functional as F
    B, H, W, Cin, Cout = 1, 4, 128, 128, 128
    w32 = randn(Cout, Cin, 3, 3, seed=1, scale=1 / math.sqrt(Cin * 9))
    x32 = randn(B, H, W, Cin, seed=2)
    xp = F.pad(F.pad(x32.permute(0, 3, 1, 2).double(), (1, 1, 0, 0), mode="circular"), (0, 0, 1, 1))
    ref = F.conv2d(xp, w32.double()).permute(0, 2, 3, 1).float()
    errs = {}
    for parts in (2, 1, 3):
        h = Both()
        w = h.t(w32)
        a = h.t(split(x32, parts))
        wp = h.t(wp_zeros(Cout, Cin, 9, parts))
        out = h.t(torch.zeros(B, H, W, Cout))
        ws = 2.0 ** ((14 if parts == 3 else 8) - math.floor(math.log2(float(w32.abs().max()))))
        h.call("pack_conv_weight", [("t", w), ("t", wp), Cout, Cin, 9, 128, 2, parts, ws])
        h.call("conv_tc", [("t", a), ("t", wp), None, None, 1.0, 1.0 / ws, ("t", out), None, B, H, W, Cin, Cout, 9, 1,
                           128, 2, parts])
        errs[parts] = rel(h.out(out)[0], ref)
    assert errs[2] < 5e-6, errs
    assert errs[3] < 5e-5, errs          # fp16 main term + fp8 correction terms: ~2e-5
    assert 5e-5 < errs[1] < 2e-3, errs


@pytest.mark.parametrize("parts", [2, 1, 3])
def test_conv_ffma_matches_contract(parts):
    h = Both()
    B, H, W, Cin, Cout, taps = 2, 4, 128, 48, 64, 9
    w = h.t(randn(Cout, Cin, 3, 3, seed=1, scale=0.05))
    a = h.t(split(randn(B, H, W, Cin, seed=2), parts))
    bias = h.t(randn(Cout, seed=3, scale=0.1))
    res = h.t(randn(B, H, W, Cout, seed=4))
    w16 = h.t(torch.zeros(Cout * Cin * taps * parts, dtype=torch.float16))
    out = h.t(torch.zeros(B, H, W, Cout))
    st = h.t(torch.zeros(B, Cout, 2, dtype=torch.float64))
    ws = 32.0 if parts < 3 else 2.0 ** 16
    h.call("pack_conv_weight_plain", [("t", w), ("t", w16), Cout, Cin, taps, parts, ws])
    g, c = h.out(w16)
    assert torch.equal(g, c)
    h.call("conv_ffma", [("t", a), ("t", w16), ("t", bias), ("t", res), 1.0, 1 / ws, ("t", out), ("t", st), B, H, W,
                         Cin, Cout, taps, 1, parts])
    g, c = h.out(out)
    assert rel(g, c) < 2e-6
    assert rel(*h.out(st)) < 1e-6


@pytest.mark.parametrize("parts", [2, 1, 3])
@pytest.mark.parametrize("C0,C1,groups,affine,ada,silu,norm", [
    (64, 0, 8, True, False, True, True),
    (64, 0, 8, False, True, True, True),
    (128, 64, 8, True, False, True, True),
    (256, 128, 32, True, True, False, True),
    (64, 64, 1, False, False, False, False),
])
def test_gn_act(parts, C0, C1, groups, affine, ada, silu, norm):
    h = Both()
    B, H, W = 3, 5, 256
    HW = H * W
    C = C0 + C1
    x0 = randn(B, HW, C0, seed=1) * 2 + 0.3
    x1 = randn(B, HW, max(C1, 8), seed=2) - 0.5
    ix0, ix1 = h.t(x0), h.t(x1)

    def stats(x):
        return torch.stack([x.double().sum(1), (x.double() ** 2).sum(1)], -1).contiguous()
    s0, s1 = h.t(stats(x0)), h.t(stats(x1[..., :max(C1, 8)]))
    gam, bet = h.t(1 + 0.1 * randn(C, seed=3)), h.t(0.1 * randn(C, seed=4))
    P = 2 * C + 24
    adat = h.t(0.3 * randn(B, P, seed=5))
    y = h.t(operand_zeros(parts, B, H, W, C))
    yr = h.t(operand_zeros(parts, B, H, W, C))
    # ada pointer offset of 8 floats inside the row, as the planner does
    args = [("t", ix0), C0, ("t", ix1) if C1 else None, C1, ("t", s0) if norm else None,
            ("t", s1) if (norm and C1) else None, ("t", gam) if affine else None, ("t", bet) if affine else None,
            ("t", adat) if ada else None, P, groups, 1e-6, 1 if silu else 0, ("t", y), ("t", yr) if C1 else None, parts,
            B, H, W]
    h.call("gn_act_f16", args)
    operand_close(*h.out(y), parts, B, H, W, C)
    if C1:
        g, c = h.out(yr)
        assert torch.equal(g, c)           # raw operand: exact encoding of the concatenated input


def test_channel_stats_and_fir():
    for up in (0, 1):
        h = Both()
        B, H, W, C = 2, 6, 32, 64
        x = h.t(randn(B, H, W, C, seed=1))
        Ho, Wo = (2 * H, 2 * W) if up else (H // 2, W // 2)
        y = h.t(torch.zeros(B, Ho, Wo, C))
        st = h.t(torch.zeros(B, C, 2, dtype=torch.float64))
        h.call("fir_resample", [("t", x), ("t", y), ("t", st), B, H, W, C, up, 1])
        assert rel(*h.out(y)) < 1e-6
        assert rel(*h.out(st)) < 1e-6
    h = Both()
    x = h.t(randn(2, 700, 128, seed=3))
    st = h.t(torch.zeros(2, 128, 2, dtype=torch.float64))
    h.call("channel_stats", [("t", x), ("t", st), 2, 700, 128])
    assert rel(*h.out(st)) < 1e-6


@pytest.mark.parametrize("parts", [2, 1, 3])
@pytest.mark.parametrize("ring", [1, 0])
def test_fir_up_operand(parts, ring):
    h = Both()
    B, H, W, C = 2, 3, 128, 64
    x = h.t(randn(B, H, W, C, seed=1))
    y = h.t(operand_zeros(parts, B, 2 * H, 2 * W, C))
    h.call("fir_up_operand", [("t", x), ("t", y), parts, B, H, W, C, ring])
    g, c = h.out(y)
    if parts == 2:
        pg, pc = load_operand(g.data_ptr(), 2, B, 2 * H, 2 * W, C), load_operand(c.data_ptr(), 2, B, 2 * H, 2 * W, C)
        assert rel(pg[0] + pg[1], pc[0] + pc[1]) < 1e-6
    operand_close(g, c, parts, B, 2 * H, 2 * W, C)


def test_time_embed():
    h = Both()
    B, Cs, E, P = 4, 64, 256, 640
    t = h.t(torch.tensor([-14.5, -3.0, 0.7, 12.0]))
    w1, b1 = h.t(randn(E, Cs, seed=1, scale=0.1)), h.t(randn(E, seed=2, scale=0.1))
    w2, b2 = h.t(randn(E, E, seed=3, scale=0.06)), h.t(randn(E, seed=4, scale=0.1))
    wp, bp = h.t(randn(P, E, seed=5, scale=0.06)), h.t(randn(P, seed=6, scale=0.1))
    add = h.t(randn(B, E, seed=7))
    temb, ada = h.t(torch.zeros(B, E)), h.t(torch.zeros(B, P))
    h.call("time_embed", [("t", t), ("t", w1), ("t", b1), ("t", w2), ("t", b2), ("t", add), ("t", wp), ("t", bp),
                          ("t", temb), ("t", ada), B, Cs, E, P])
    assert rel(*h.out(temb)) < 1e-5
    assert rel(*h.out(ada)) < 1e-5


def test_in_conv_out_conv_direct():
    B, H, Cx, Cout = 2, 5, 2, 64
    for W, ring in ((64, 1), (256, 1), (128, 0)):        # generic kernel / row-tiled kernel (W % 128 == 0)
        h = Both()
        x = h.t(randn(B, Cx, H, W, seed=1))
        w = h.t(randn(Cout, Cx, 3, 3, seed=2, scale=0.2))
        cst = h.t(randn(1, H, W, Cout, seed=3))
        out = h.t(torch.zeros(B, H, W, Cout))
        st = h.t(torch.zeros(B, Cout, 2, dtype=torch.float64))
        h.call("in_conv", [("t", x), ("t", w), ("t", cst), 0, ("t", out), ("t", st), B, H, W, Cx, Cout, ring])
        assert rel(*h.out(out)) < 1e-6
        assert rel(*h.out(st)) < 1e-6
    W = 64

    h = Both()
    Cin, Co = 30, 64
    x = h.t(randn(1, H, W, Cin, seed=1))
    w = h.t(randn(Co, Cin, 3, 3, seed=2, scale=0.1))
    b = h.t(randn(Co, seed=3))
    out = h.t(torch.zeros(1, H, W, Co))
    h.call("conv_direct_f32", [("t", x), ("t", w), ("t", b), ("t", out), 1, H, W, Cin, Co, 3, 1])
    assert rel(*h.out(out)) < 1e-6

    for is16 in (0, 1):
        h = Both()
        a32 = randn(B, H, W, 64, seed=4)
        a = h.t(a32.half() if is16 else a32)
        w = h.t(randn(2, 64, 3, 3, seed=5, scale=0.1))
        b = h.t(randn(2, seed=6))
        pred = h.t(torch.zeros(B, 2, H, W))
        h.call("out_conv", [("t", a), is16, ("t", w), ("t", b), ("t", pred), B, H, W, 64, 2, 1])
        assert rel(*h.out(pred)) < 1e-6


@pytest.mark.parametrize("mode,objective", [(0, 0), (0, 1), (0, 2), (1, 0)])
def test_sampler_update(mode, objective):
    h = Both()
    B, n = 3, 5000
    xt, pr, nz = h.t(randn(B, n, seed=1)), h.t(randn(B, n, seed=2)), h.t(randn(B, n, seed=3))
    lt, ls = torch.tensor([-4.0, 0.5, 3.0]), torch.tensor([-3.0, 1.5, 5.0])
    a_t, s_t, a_s, s_s = lt.sigmoid().sqrt(), (-lt).sigmoid().sqrt(), ls.sigmoid().sqrt(), (-ls).sigmoid().sqrt()
    c1 = 0.3 * s_s / s_t * (1 - a_t ** 2 / a_s ** 2).sqrt()
    c2 = (1 - a_s ** 2 - c1 ** 2).sqrt()
    cc = -torch.expm1(lt - ls)
    coef = h.t(torch.stack([a_t, s_t, a_s, s_s, c1, c2, cc, torch.zeros(3)], -1))
    xs = h.t(torch.zeros(B, n))
    h.call("sampler_update", [("t", xt), ("t", pr), ("t", nz), ("t", coef), ("t", xs), B, n, mode, objective, 1.0])
    assert rel(*h.out(xs)) < 1e-6


def test_gn_act_f32():
    h = Both()
    B, HW, C = 2, 300, 128
    x = randn(B, HW, C, seed=1) * 1.7 + 0.2
    ix = h.t(x)
    st = h.t(torch.stack([x.double().sum(1), (x.double() ** 2).sum(1)], -1).contiguous())
    gam, bet = h.t(1 + 0.1 * randn(C, seed=3)), h.t(0.1 * randn(C, seed=4))
    y = h.t(torch.zeros(B, HW, C))
    h.call("gn_act_f32", [("t", ix), ("t", st), ("t", gam), ("t", bet), 32, 1e-5, 1, ("t", y), B, HW, C])
    assert rel(*h.out(y)) < 3e-6


@pytest.mark.parametrize("parts", [2, 1, 3])
@pytest.mark.parametrize("E,heads,T,W", [(512, 8, 512, 128), (256, 8, 512, 128), (256, 8, 128, 128), (512, 8, 384, 128)])
def test_flash_attention(parts, E, heads, T, W):
    h = Both()
    B = 2
    d = E // heads
    qkv = h.t(randn(B, T, 3 * E, seed=1))
    out = h.t(operand_zeros(parts, B, T // W, W, E))
    ws = h.t(torch.zeros(h.gpu.flash_attention_workspace(B, heads, T, 0, d, d), dtype=torch.uint8))
    h.call("flash_attention", [("t", qkv), E, ("t", out), W, parts, B, heads, T, 1 / math.sqrt(d), ("t", ws)])
    operand_close(*h.out(out), parts, B, T // W, W, E)


@pytest.mark.parametrize("parts", [2, 1, 3])
@pytest.mark.parametrize("C,T,W", [(256, 2048, 256), (512, 512, 128)])
def test_flash_attention_oa(parts, C, T, W):
    h = Both()
    B, L2 = 2, 13
    qkv = h.t(randn(B, T, 3 * C, seed=1))
    pos_p = h.t(randn(B, T, C, seed=2))
    kl, pos_l, vl = h.t(randn(B, L2, C, seed=3)), h.t(randn(B, L2, C, seed=4)), h.t(randn(B, L2, C, seed=5))
    out = h.t(operand_zeros(parts, B, T // W, W, C))
    ws = h.t(torch.zeros(h.gpu.flash_attention_workspace(B, C // 32, T, L2, 64, 32), dtype=torch.uint8))
    h.call("flash_attention_oa", [("t", qkv), ("t", pos_p), ("t", kl), ("t", pos_l), ("t", vl), ("t", out), W, parts, B,
                                  C, C // 32, T, L2, 1 / math.sqrt(64), ("t", ws)])
    operand_close(*h.out(out), parts, B, T // W, W, C)


@pytest.mark.parametrize("ring", [1, 0])
def test_out_conv_row_tiled(ring):
    """W % 128 == 0 fp32 input takes the row-tiled kernel (one read of every input pixel)."""
    h = Both()
    B, H, W = 2, 6, 256
    a = h.t(randn(B, H, W, 64, seed=4))
    w = h.t(randn(2, 64, 3, 3, seed=5, scale=0.1))
    b = h.t(randn(2, seed=6))
    pred = h.t(torch.zeros(B, 2, H, W))
    h.call("out_conv", [("t", a), 0, ("t", w), ("t", b), ("t", pred), B, H, W, 64, 2, ring])
    assert rel(*h.out(pred)) < 1e-6
