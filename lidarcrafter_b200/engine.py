"""Host-side execution plans: turn a denoiser module (reference-compatible parameters) into a static
sequence of libb200lidar kernel launches over pre-allocated NHWC device buffers.

Layout in HBM: residual stream / conv outputs fp32 [B, H*W, C]; conv operands fp16
[B, H*W, C]; GroupNorm statistics fp64 [B, C, 2] (sum, sum of squares) in one arena that is zeroed once
per forward; weights repacked once into the tensor-core tile image (fp16).
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass

import torch

from . import _lib

NUM_SMS = 148


def _ptr(t):
    return 0 if t is None else t.data_ptr()


@dataclass
class Act:
    """An NHWC activation: fp32 tensor + (optional) per-channel statistics slice."""
    t: torch.Tensor            # [B, H*W, C] fp32
    H: int
    W: int
    C: int
    stats: dict | None = None  # stats slot from Plan.new_stats(): [B, C, 2] fp64 view into the arena


def conv_merged(bn: int, rows: int, parts: int) -> bool:
    """mirror of b200_conv_merged(): hi/lo weight rows adjacent when 2 x rows x bn accumulator columns fit twice in TMEM"""
    return parts == 2 and 2 * rows * bn <= 256


def pick_tile(B: int, H: int, W: int, Cout: int, taps: int, parts: int = 2, sms: int = NUM_SMS):
    """Choose (bn, rows) for b200_conv_tc: maximise (SM fill) x min(L2-feed, MMA-shape) efficiency."""
    best = None
    for bn in (128, 64):
        if Cout % bn:
            continue
        for rows in (4, 2, 1):
            if H % rows or rows * bn > 256:
                continue
            tiles = B * (H // rows) * (W // 128) * (Cout // bn)
            waves = math.ceil(tiles / sms)
            fill = tiles / (waves * sms)
            halo = 2 if taps == 9 else 0
            l2 = 64.0 / rows + 924.0 * (rows + halo) / (rows * bn)      # bytes / clk / SM needed from L2
            feed = min(1.0, 40.0 / l2)
            shape = 1.0 if bn == 128 else 0.7
            score = fill * min(feed, shape)
            if best is None or score > best[0] + 1e-9:
                best = (score, bn, rows)
    if best is None:
        raise ValueError(f"no conv tile for Cout={Cout}, H={H}")
    return best[1], best[2]


def weight_scale(weight: torch.Tensor, parts: int = 2) -> float:
    """Power of two s, undone exactly in the conv epilogue (w_inv = 1 / s).
    parts 1/2: max|w| * s in [256, 512) -- keeps the fp16 hi/lo split of small weights out of the subnormal range.
    parts 3 (fp16 + fp8 correction): max|w| * s in [2^14, 2^15) so that BOTH fp8 correction operands, e4m3(w s 2^-11)
    (max in [8, 16)) and e4m3(w s - fp16(w s)) (|.| <= 8), sit in the normal e4m3 range."""
    m = float(weight.detach().abs().max())
    if not (m > 0.0) or not math.isfinite(m):
        return 1.0
    return float(2.0 ** ((14 if parts >= 3 else 8) - math.floor(math.log2(m))))


# conv precision modes (the ``precision`` attribute of the UNet mirrors) -> ``parts`` of the C-ABI
#   fp16x3: error-compensated fp16 split, 3 fp16 MMAs / product, ~2e-6 relative through the UNet (fp32-grade)
#   fp16f8: fp16 main term + ONE fp8 (e4m3) MMA for both correction cross terms, ~5e-5 relative (tolerance 1e-3)
#   fp16  : single fp16 pass, ~1.7e-3 (outside the 1e-3 tolerance; kept for measurement only)
PRECISION_PARTS = {"fp16x3": 2, "fp16f8": 3, "fp16": 1}


def precision_parts(precision: str) -> int:
    if precision not in PRECISION_PARTS:
        raise ValueError(f"precision must be one of {sorted(PRECISION_PARTS)}, got {precision!r}")
    return PRECISION_PARTS[precision]


def planes(parts: int) -> int:
    """fp16-sized planes of a conv operand: parts 1 -> 1; 2 (hi, lo) -> 2; 3 (fp16 hi + e4m3 pair) -> 2"""
    return 1 if parts == 1 else 2


_TUNE_CACHE: dict = {}
_TUNE_MS: dict = {}          # measured ms per launch of the chosen tile (this process only)
_TUNE_FILE_LOADED = False


def _tune_file():
    """B200_TUNE_FILE=path.json: persist the measured (bn, rows) choices, so that a second process (e.g. the same bench
    under ncu, where timing launches is distorted by the profiler) runs exactly the tiles the first one measured."""
    return os.environ.get("B200_TUNE_FILE", "")


def _load_tune_file():
    global _TUNE_FILE_LOADED
    if _TUNE_FILE_LOADED:
        return
    _TUNE_FILE_LOADED = True
    path = _tune_file()
    if path and os.path.exists(path):
        import json
        for k, v in json.load(open(path)).items():
            _TUNE_CACHE[tuple(json.loads(k))] = tuple(v)


def _save_tune_file():
    path = _tune_file()
    if path:
        import json
        with open(path, "w") as f:
            json.dump({json.dumps([x if isinstance(x, (bool, str)) else int(x) for x in k]): list(v)
                       for k, v in _TUNE_CACHE.items()}, f)


def tile_candidates(B: int, H: int, W: int, Cout: int, parts: int):
    out = []
    for bn in (128, 64):
        if Cout % bn:
            continue
        for rows in (4, 2, 1):
            if H % rows == 0 and rows * bn <= 256:
                out.append((bn, rows))
    return out


def col_walk_ok(B: int, H: int, W: int, C0: int, C1: int, Cout: int, taps: int, parts: int, sms: int = NUM_SMS) -> bool:
    """shapes the column-walk variant of the fused conv accepts (csrc/conv_col.cuh, tile code rows = 0): 64 -> 64 channels,
    3x3, fp16f8 operands, and a CTA's run of row tiles inside at most two samples"""
    if not (parts == 3 and taps == 9 and C0 == 64 and C1 == 0 and Cout == 64 and W % 128 == 0):
        return False
    n = B * H * (W // 128)
    grid = min(n, sms)
    return math.ceil(n / grid) <= H * (W // 128)


class PackedConv:
    """fp16 tile image of one conv's weights + fp32 bias (device)."""

    def __init__(self, lib, weight: torch.Tensor, bias: torch.Tensor | None, bn: int, rows: int, parts: int, stream: int):
        Cout, Cin, kh, kw = weight.shape
        self.Cout, self.Cin, self.taps, self.bn, self.rows, self.parts = Cout, Cin, kh * kw, bn, rows, parts
        w = weight.detach().float().contiguous()
        self.wscale = weight_scale(w, parts)
        self.packed = torch.empty(Cout * Cin * self.taps * planes(parts), dtype=torch.float16, device=w.device)
        lib.pack_conv_weight(_ptr(w), _ptr(self.packed), Cout, Cin, self.taps, bn, rows, parts, self.wscale, stream)
        self.bias = None if bias is None else bias.detach().float().contiguous()
        self._keep = w


class Plan:
    """Recorded launch list for one (module, batch) pair.  ``run(stream)`` replays it."""

    def __init__(self, lib, device, B: int, conv_impl: str = "tc", parts: int = 2):
        self.lib = lib
        self.device = device
        self.B = B
        # 2 = error-compensated fp16 split (fp32-grade), 3 = fp16 + fp8 correction terms (~5e-5), 1 = single fp16 pass
        self.parts = parts
        self.ops = []                 # list of (callable, args-without-stream)
        self.meta = []                # (kernel name, algorithmic flops, algorithmic bytes) per op
        self.tags = []                # per op: the block it belongs to (PlanBuilder.tag; "" outside a tagged block)
        self.tag = ""
        self.bufs = []                # keep-alive
        self.stats_chunks = []
        self.stats_arena = None
        self.conv_impl = conv_impl
        self.flops = 0.0              # algorithmic conv / attention FLOPs per forward (for reporting)

    # ---- buffers ----
    def f32(self, *shape):
        t = torch.empty(*shape, dtype=torch.float32, device=self.device)
        self.bufs.append(t)
        return t

    def operand(self, H: int, W: int, C: int):
        """conv operand buffer, layout [planes][B][H][W/128][C/8][130][8] (fp16-sized elements; 128-pixel tiles with
        their two ring-halo pixels -- csrc/common.cuh; plane 1 of parts 3 holds e4m3 pairs)"""
        if W % 128 or C % 8:
            raise ValueError(f"conv operands need W % 128 == 0 and C % 8 == 0 (got W={W}, C={C})")
        t = torch.empty(planes(self.parts), self.B * H * (W // 128) * (C // 8) * 130 * 8, dtype=torch.float16,
                        device=self.device)
        self.bufs.append(t)
        return t

    def new_stats(self, C: int):
        """Reserve a [B, C, 2] fp64 slot; materialised by finalize()."""
        off = sum(n for _, n in self.stats_chunks)
        holder = {"off": off, "C": C, "t": None}
        self.stats_chunks.append((holder, self.B * C * 2))
        return holder

    def finalize(self):
        total = sum(n for _, n in self.stats_chunks)
        self.stats_arena = torch.zeros(max(total, 1), dtype=torch.float64, device=self.device)
        for holder, n in self.stats_chunks:
            holder["t"] = self.stats_arena[holder["off"]:holder["off"] + n]
        # resolve late-bound pointers once
        self.ops = [(fn, tuple(a() if callable(a) else a for a in args)) for fn, args in self.ops]

    def add(self, fn, *args, name: str = "", flops: float = 0.0, nbytes: float = 0.0):
        self.ops.append((fn, args))
        self.meta.append((name or getattr(fn, "__name__", "op"), flops, nbytes))
        self.tags.append(self.tag)

    def profile(self, stream: int, reps: int = 3):
        """Per-launch device time of every op INSIDE a replay of the whole step: CUDA events on the launching stream between
        consecutive launches, the whole sequence queued behind a ~1.5 ms spin kernel so that the host stays ahead of the GPU.
        The interval of a launch is then the kernel as it runs in the step -- same order, same L2 contents (its inputs were
        written by the launch before it) -- plus one event; timing every op on its own long after its producer (the earlier
        method) read every input cold from HBM.  Best of ``reps`` replays per op.
        Returns [(name, ms, algorithmic flops, algorithmic bytes)]."""
        self.run(stream)
        torch.cuda.synchronize(self.device)
        n = len(self.ops)
        best = [float("inf")] * n
        for _ in range(reps):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(n + 1)]
            with torch.inference_mode():
                self.stats_arena.zero_()
            torch.cuda._sleep(3_000_000)
            ev[0].record()
            for i, (fn, args) in enumerate(self.ops):
                fn(*args, stream)
                ev[i + 1].record()
            ev[n].synchronize()
            for i in range(n):
                best[i] = min(best[i], ev[i].elapsed_time(ev[i + 1]))
        return [(name, best[i], fl, by) for i, (name, fl, by) in enumerate(self.meta)]

    def run(self, stream: int):
        with torch.inference_mode():     # buffers may have been created under sample()'s inference_mode
            self.stats_arena.zero_()
        for fn, args in self.ops:
            fn(*args, stream)

    @property
    def n_kernels(self):
        # + the arena memset; a split-K conv is two launches (K slices + reduce / epilogue)
        return len(self.ops) + 1 + sum(1 for name, _, _ in self.meta if name == "conv_tc_splitk")


def _sp(holder):
    """late-bound pointer of a stats slot (arena is allocated in finalize())."""
    if holder is None:
        return 0
    return lambda: holder["t"].data_ptr()


class PlanBuilder:
    """Shared building blocks used by the EfficientUNet / LayoutUnetV1 planners."""

    def __init__(self, plan: Plan, ring: bool, stream: int):
        self.p = plan
        self.lib = plan.lib
        self.ring = 1 if ring else 0
        self.stream = stream
        self.B = plan.B
        # GroupNorm(+AdaGN)+SiLU+operand split fused IN FRONT of the conv (b200_conv_gn_tc) or as a separate gn_act launch +
        # operand round trip: "auto" (default) measures both per layer shape at plan build and keeps the faster one -- the
        # fused front end re-does the elementwise work per halo row and per output-channel tile, so it wins where launches
        # are short (small batch) or the layer has ONE n-tile, and loses on the wide layers (profiles/r02_fused_front_*);
        # B200_FUSE_FRONT=1 / 0 force one of them (A/B measurements, tests).
        self.fuse_front = os.environ.get("B200_FUSE_FRONT", "auto")

    # ---- conv on tensor cores (or the FFMA cross-check path) ----
    def conv(self, a16: torch.Tensor, H: int, W: int, weight, bias, res: torch.Tensor | None, scale: float,
             want_stats: bool):
        """conv of a pre-built operand (attention output, FIR-upsampled operand) -> (out, stats slot)"""
        Cout, Cin, kh, kw = weight.shape
        taps = kh * kw
        out = self.p.f32(self.B, H * W, Cout)
        st = self.p.new_stats(Cout) if want_stats else None
        fl = 2.0 * self.B * H * W * taps * Cin * Cout
        self.p.flops += fl
        npix = self.B * H * W
        by = npix * (2.0 * planes(self.p.parts) * Cin + 4.0 * Cout * (2 if res is not None else 1)) \
            + 2.0 * planes(self.p.parts) * taps * Cin * Cout
        if self.p.conv_impl == "tc":
            bn, rows = self.tune_tile(a16, H, W, weight, bias, res, out)
            pc = PackedConv(self.lib, weight, bias, bn, rows, self.p.parts, self.stream)
            self.p.bufs.append(pc)
            splits = self.tune_split(a16, H, W, pc, res, out, bn, rows)
            if splits > 1:
                ws = self.p.f32(splits, self.B * H * W * Cout)
                self.p.add(self.lib.conv_tc_splitk, _ptr(a16), _ptr(pc.packed), _ptr(pc.bias), _ptr(res), float(scale),
                           1.0 / pc.wscale, _ptr(out), _sp(st), _ptr(ws), splits, self.B, H, W, Cin, Cout, taps, self.ring, bn,
                           rows, self.p.parts, name="conv_tc_splitk", flops=fl, nbytes=by)
            else:
                self.p.add(self.lib.conv_tc, _ptr(a16), _ptr(pc.packed), _ptr(pc.bias), _ptr(res), float(scale),
                           1.0 / pc.wscale, _ptr(out), _sp(st), self.B, H, W, Cin, Cout, taps, self.ring, bn, rows,
                           self.p.parts, name="conv_tc", flops=fl, nbytes=by)
        else:
            w = weight.detach().float().contiguous()
            ws = weight_scale(w, self.p.parts)
            w16 = torch.empty(Cout * Cin * taps * self.p.parts, dtype=torch.float16, device=w.device)
            self.lib.pack_conv_weight_plain(_ptr(w), _ptr(w16), Cout, Cin, taps, self.p.parts, ws, self.stream)
            b32 = None if bias is None else bias.detach().float().contiguous()
            self.p.bufs += [w, w16, b32]
            self.p.add(self.lib.conv_ffma, _ptr(a16), _ptr(w16), _ptr(b32), _ptr(res), float(scale), 1.0 / ws,
                       _ptr(out), _sp(st), self.B, H, W, Cin, Cout, taps, self.ring, self.p.parts,
                       name="conv_ffma", flops=fl, nbytes=by)
        return out, st

    def conv_gn(self, srcs: list[Act], weight, bias, res: torch.Tensor | None, scale: float, want_stats: bool,
                gamma=None, beta=None, groups: int = 1, eps: float = 0.0, silu: bool = False, ada=None, ada_stride: int = 0,
                ada_off: int = 0, normalize: bool = True):
        """conv(act(GroupNorm[+AdaGN](cat(srcs)))) -> (out, stats slot) in ONE launch (b200_conv_gn_tc): the transform warps
        of the conv kernel normalise / activate / split the fp32 activation(s) straight into the tensor-core operand slab.
        ``normalize=False``: plain fp32 -> operand conversion (1x1 skip conv of a ResBlock, conv in front of a down-sampler).
        The cross-check paths (conv_impl "ffma", single-pass fp16, B200_FUSE_FRONT=0) run gn_act + conv instead."""
        a0 = srcs[0]
        a1 = srcs[1] if len(srcs) > 1 else None
        H, W = a0.H, a0.W
        Cout, Cin, kh, kw = weight.shape
        assert Cin == a0.C + (a1.C if a1 else 0)
        fused = (self.p.conv_impl == "tc" and self.fuse_front != "0" and self.p.parts in (2, 3) and a0.C % 16 == 0
                 and (a1 is None or a1.C % 16 == 0) and Cin <= 1024 and groups <= 32)
        if fused and self.fuse_front == "auto" and self.p.device.type == "cuda":
            fused = self.tune_front(srcs, weight, bias, res, silu)
        if not fused:
            a16 = self.gn_act(srcs, gamma, beta, groups, eps, silu, ada=ada, ada_stride=ada_stride, ada_off=ada_off,
                              normalize=normalize)
            return self.conv(a16, H, W, weight, bias, res, scale, want_stats)
        if normalize:
            assert a0.stats is not None and (a1 is None or a1.stats is not None)
        taps = kh * kw
        out = self.p.f32(self.B, H * W, Cout)
        st = self.p.new_stats(Cout) if want_stats else None
        fl = 2.0 * self.B * H * W * taps * Cin * Cout
        self.p.flops += fl
        npix = self.B * H * W
        by = npix * (4.0 * Cin + 4.0 * Cout * (2 if res is not None else 1)) + 2.0 * planes(self.p.parts) * taps * Cin * Cout
        g = None if gamma is None else gamma.detach().float().contiguous()
        b = None if beta is None else beta.detach().float().contiguous()
        self.p.bufs += [g, b]
        ada_ptr = 0 if ada is None else ada.data_ptr() + 4 * ada_off
        front = (_ptr(a0.t), a0.C, _ptr(a1.t) if a1 else 0, a1.C if a1 else 0,
                 _sp(a0.stats) if normalize else 0, _sp(a1.stats) if (normalize and a1) else 0, _ptr(g), _ptr(b), ada_ptr,
                 ada_stride, groups, float(eps), 1 if silu else 0)
        bn, rows = self.tune_tile(None, H, W, weight, bias, res, out, front=(_ptr(a0.t), a0.C, _ptr(a1.t) if a1 else 0,
                                                                             a1.C if a1 else 0))
        pc = PackedConv(self.lib, weight, bias, bn, rows, self.p.parts, self.stream)
        self.p.bufs.append(pc)
        self.p.add(self.lib.conv_gn_tc, *front, _ptr(pc.packed), _ptr(pc.bias), _ptr(res), float(scale), 1.0 / pc.wscale,
                   _ptr(out), _sp(st), self.B, H, W, Cout, taps, self.ring, bn, rows, self.p.parts, name="conv_gn_tc",
                   flops=fl, nbytes=by)
        return out, st

    def tune_front(self, srcs: list[Act], weight, bias, res, silu: bool) -> bool:
        """True if the fused front end (one b200_conv_gn_tc launch) is faster than gn_act_f16 + b200_conv_tc for this layer
        shape: both are timed on the device with their own best tile (CUDA events), cached per shape for the process."""
        a0 = srcs[0]
        a1 = srcs[1] if len(srcs) > 1 else None
        H, W = a0.H, a0.W
        Cout, Cin, kh, kw = weight.shape
        key = ("front", self.B, H, W, a0.C, a1.C if a1 else 0, Cout, kh * kw, self.p.parts, res is not None, bool(silu))
        _load_tune_file()
        if key in _TUNE_CACHE:
            return bool(_TUNE_CACHE[key][0])
        out = torch.empty(self.B, H * W, Cout, dtype=torch.float32, device=self.p.device)
        y = torch.empty(planes(self.p.parts), self.B * H * (W // 128) * (Cin // 8) * 130 * 8, dtype=torch.float16,
                        device=self.p.device)
        front = (_ptr(a0.t), a0.C, _ptr(a1.t) if a1 else 0, a1.C if a1 else 0)
        for fr in (True, False):     # (tiles restored from B200_TUNE_FILE carry no time: measure again)
            k = (self.B, H, W, Cin, Cout, kh * kw, self.p.parts, res is not None, fr)
            if k in _TUNE_CACHE and k not in _TUNE_MS:
                del _TUNE_CACHE[k]
        self.tune_tile(None, H, W, weight, bias, res, out, front=front)
        self.tune_tile(y, H, W, weight, bias, res, out)
        gn_args = tuple(front) + (0, 0, 0, 0, 0, 0, 1, 0.0, 1 if silu else 0, _ptr(y), 0, self.p.parts, self.B, H, W, self.stream)
        t_gn = self._time(self.lib.gn_act_f16, gn_args)
        t_fused = _TUNE_MS[(self.B, H, W, Cin, Cout, kh * kw, self.p.parts, res is not None, True)]
        t_conv = _TUNE_MS[(self.B, H, W, Cin, Cout, kh * kw, self.p.parts, res is not None, False)]
        _TUNE_CACHE[key] = (1 if t_fused < t_gn + t_conv else 0, 0)
        _save_tune_file()
        return bool(_TUNE_CACHE[key][0])

    def tune_split(self, a16, H, W, pc, res, out, bn: int, rows: int) -> int:
        """K slices per output tile for b200_conv_tc_splitk (1 = plain b200_conv_tc): only layers whose tiles fill less than half
        of the SMs are candidates (deep levels at small batch: few tiles, long K, every CTA streams the whole weight tile);
        conv + reduce kernel are timed against the plain conv on the device, cached per shape.  B200_SPLIT_K=0 disables."""
        Cout, Cin, taps = pc.Cout, pc.Cin, pc.taps
        tiles = self.B * (H // rows) * (W // 128) * (Cout // bn)
        nch = Cin // (32 if self.p.parts == 1 else 16)
        key = ("split", self.B, H, W, Cin, Cout, taps, self.p.parts, res is not None, bn, rows)
        _load_tune_file()
        if key in _TUNE_CACHE:
            return int(_TUNE_CACHE[key][0])
        if (self.p.device.type != "cuda" or os.environ.get("B200_SPLIT_K", "1") == "0" or 2 * tiles > NUM_SMS
                or Cout // 4 > 256 or 256 % (Cout // 4)):
            return 1
        tail = (_ptr(pc.packed), _ptr(pc.bias), _ptr(res), 1.0, 1.0 / pc.wscale, _ptr(out), 0)
        dims = (self.B, H, W, Cin, Cout, taps, self.ring, bn, rows, self.p.parts, self.stream)
        best = (self._time(self.lib.conv_tc, (_ptr(a16),) + tail + dims), 1)
        for s in (2, 4, 8):
            if nch % s or tiles * s > 2 * NUM_SMS:
                continue
            ws = torch.empty(s, self.B * H * W * Cout, dtype=torch.float32, device=self.p.device)
            ms = self._time(self.lib.conv_tc_splitk, (_ptr(a16),) + tail + (_ptr(ws), s) + dims)
            if ms < best[0]:
                best = (ms, s)
        _TUNE_CACHE[key] = (best[1], 0)
        _save_tune_file()
        return best[1]

    @staticmethod
    def _time(fn, args, batches: int = 3, n: int = 5) -> float:
        """ms per launch: best of `batches` batches of `n` back-to-back launches (one batch lets power-cap / clock noise decide;
        single launches are dominated by the launch gap, and by the interception cost under a profiler)"""
        for _ in range(2):                           # warm-up (function attributes, caches, clocks)
            fn(*args)
        ms = float("inf")
        for _ in range(batches):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn(*args)
            e1.record()
            e1.synchronize()
            ms = min(ms, e0.elapsed_time(e1) / n)
        return ms

    def tune_tile(self, a16, H, W, weight, bias, res, out, front=None):
        """(bn, rows) of b200_conv_tc / b200_conv_gn_tc for this shape: measured once per shape on the device (CUDA events,
        best of the candidate tiles), cached for the process; the analytic pick_tile() is only the no-GPU fallback of the
        planner (emulator tests).  ``front`` = (x0, C0, x1, C1) of the fused front end (timed with SiLU, no statistics)."""
        Cout, Cin, kh, kw = weight.shape
        taps = kh * kw
        key = (self.B, H, W, Cin, Cout, taps, self.p.parts, res is not None, front is not None)
        _load_tune_file()
        if key in _TUNE_CACHE:
            return _TUNE_CACHE[key]
        if self.p.device.type != "cuda":
            return pick_tile(self.B, H, W, Cout, taps, self.p.parts)
        best = None
        packed = {}
        cands = tile_candidates(self.B, H, W, Cout, self.p.parts)
        if front is not None and col_walk_ok(self.B, H, W, front[1], front[3], Cout, taps, self.p.parts) \
                and os.environ.get("B200_COL_WALK", "1") != "0":
            cands.append((64, 0))     # column walk: every input row converted once, all weights resident (csrc/conv_col.cuh)
        for bn, rows in cands:
            if front is not None and taps == 9 and rows == 4:
                continue          # 6 staged rows: the transform warps' register-resident prefetch stage does not fit (spills)
            pk = (bn, conv_merged(bn, rows, self.p.parts), rows == 0)
            if pk not in packed:
                packed[pk] = PackedConv(self.lib, weight, bias, bn, rows, self.p.parts, self.stream)
            pc = packed[pk]
            tail = (_ptr(pc.packed), _ptr(pc.bias), _ptr(res), 1.0, 1.0 / pc.wscale, _ptr(out), 0, self.B, H, W)
            if front is None:
                fn = self.lib.conv_tc
                args = (_ptr(a16),) + tail + (Cin, Cout, taps, self.ring, bn, rows, self.p.parts, self.stream)
            else:
                fn = self.lib.conv_gn_tc
                args = tuple(front) + (0, 0, 0, 0, 0, 0, 1, 0.0, 1) + tail + (Cout, taps, self.ring, bn, rows, self.p.parts,
                                                                             self.stream)
            ms = self._time(fn, args)
            if best is None or ms < best[0]:
                best = (ms, bn, rows)
        _TUNE_CACHE[key] = (best[1], best[2])
        _TUNE_MS[key] = best[0]
        _save_tune_file()
        return _TUNE_CACHE[key]

    # ---- GroupNorm(+AdaGN)+SiLU -> fp16 operand ----
    def gn_act(self, srcs: list[Act], gamma, beta, groups: int, eps: float, silu: bool, ada=None, ada_stride=0,
               ada_off=0, normalize: bool = True, also_raw: bool = False):
        """-> fp16 conv operand; with ``also_raw`` additionally the un-normalised input as a second operand
        (returns (y, y_raw)): one pass over x feeds norm1->conv1 and the 1x1 skip conv."""
        a0 = srcs[0]
        a1 = srcs[1] if len(srcs) > 1 else None
        C = a0.C + (a1.C if a1 else 0)
        HW = a0.H * a0.W
        y = self.p.operand(a0.H, a0.W, C)
        y_raw = self.p.operand(a0.H, a0.W, C) if also_raw else None
        g = None if gamma is None else gamma.detach().float().contiguous()
        b = None if beta is None else beta.detach().float().contiguous()
        self.p.bufs += [g, b]
        if normalize:
            assert a0.stats is not None and (a1 is None or a1.stats is not None)
        ada_ptr = 0 if ada is None else ada.data_ptr() + 4 * ada_off
        self.p.add(self.lib.gn_act_f16, _ptr(a0.t), a0.C, _ptr(a1.t) if a1 else 0, a1.C if a1 else 0,
                   _sp(a0.stats) if normalize else 0, _sp(a1.stats) if (normalize and a1) else 0, _ptr(g), _ptr(b),
                   ada_ptr, ada_stride, groups, float(eps), 1 if silu else 0, _ptr(y), _ptr(y_raw), self.p.parts, self.B,
                   a0.H, a0.W, name="gn_act_f16",
                   nbytes=self.B * HW * C * (4.0 + 2.0 * planes(self.p.parts) * (2 if also_raw else 1)))
        return (y, y_raw) if also_raw else y

    def cast16(self, srcs: list[Act]) -> torch.Tensor:
        return self.gn_act(srcs, None, None, 1, 0.0, False, normalize=False)

    def fir_up_operand(self, x: Act) -> tuple[torch.Tensor, int, int]:
        """FIR x2 upsample written straight into the operand of the conv that follows -> (operand, 2H, 2W)"""
        Ho, Wo = 2 * x.H, 2 * x.W
        y = self.p.operand(Ho, Wo, x.C)
        self.p.add(self.lib.fir_up_operand, _ptr(x.t), _ptr(y), self.p.parts, self.B, x.H, x.W, x.C, self.ring,
                   name="fir_resample", nbytes=self.B * x.C * x.H * x.W * (4.0 + 4 * 2.0 * planes(self.p.parts)))
        return y, Ho, Wo

    def fir(self, x: Act, up: bool, want_stats: bool) -> Act:
        Ho, Wo = (2 * x.H, 2 * x.W) if up else (x.H // 2, x.W // 2)
        y = self.p.f32(self.B, Ho * Wo, x.C)
        st = self.p.new_stats(x.C) if want_stats else None
        self.p.add(self.lib.fir_resample, _ptr(x.t), _ptr(y), _sp(st), self.B, x.H, x.W, x.C, 1 if up else 0, self.ring,
                   name="fir_resample", nbytes=4.0 * self.B * x.C * (x.H * x.W + Ho * Wo))
        return Act(y, Ho, Wo, x.C, st)
