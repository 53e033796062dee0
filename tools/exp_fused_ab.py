"""A/B of the fused front end: b200_conv_gn_tc (GroupNorm + SiLU + split inside the conv) against the separate launches
gn_act_f16 -> conv_tc, per layer shape of the EfficientUNet step and per candidate tile.  us per launch, CUDA events,
best of 5 x 10 launches.  PARTS=2|3 (fp16x3 | fp16f8), COUNTERS=1 adds the in-kernel wait counters of the fused kernel."""
import math
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from lidarcrafter_b200 import _lib  # noqa: E402

SHAPES = [  # B, H, W, C0, C1, Cout, taps, res
    (8, 32, 1024, 64, 0, 64, 9, 1),
    (8, 32, 1024, 64, 0, 64, 9, 0),
    (8, 32, 1024, 64, 64, 64, 9, 0),
    (8, 16, 512, 128, 0, 128, 9, 1),
    (8, 16, 512, 64, 0, 128, 9, 0),
    (8, 8, 256, 256, 0, 256, 9, 1),
    (8, 4, 128, 512, 0, 512, 9, 1),
    (8, 16, 512, 64, 0, 64, 9, 1),
    (1, 32, 1024, 64, 0, 64, 9, 1),
]


def timeit(fn, n=10, rounds=5):
    fn()
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(rounds):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        e1.synchronize()
        best = min(best, e0.elapsed_time(e1) / n * 1e3)
    return best


def main():
    lib = _lib.Lib(os.environ["LIBPATH"]) if "LIBPATH" in os.environ else _lib.get_lib()   # e.g. the -DB200_CONV_ABLATE build
    masks = [int(m) for m in os.environ.get("MASKS", "0").split(",")]
    dev = torch.device("cuda")
    st = torch.cuda.current_stream().cuda_stream
    parts = int(os.environ.get("PARTS", "2"))
    only = os.environ.get("ONLY")
    for si, (B, H, W, C0, C1, Cout, taps, res) in enumerate(SHAPES):
        if only is not None and str(si) not in only.split(","):
            continue
        Cin = C0 + C1
        kk = 3 if taps == 9 else 1
        w = (torch.randn(Cout, Cin, kk, kk, device=dev) / math.sqrt(Cin * taps)).contiguous()
        x0 = torch.randn(B, H * W, C0, device=dev)
        x1 = torch.randn(B, H * W, C1, device=dev) if C1 else None
        s0 = torch.zeros(B, C0, 2, dtype=torch.float64, device=dev)
        s1 = torch.zeros(B, C1, 2, dtype=torch.float64, device=dev) if C1 else None
        lib.channel_stats(x0.data_ptr(), s0.data_ptr(), B, H * W, C0, st)
        if C1:
            lib.channel_stats(x1.data_ptr(), s1.data_ptr(), B, H * W, C1, st)
        gam, bet = torch.ones(Cin, device=dev), torch.zeros(Cin, device=dev)
        y = torch.empty(2, B * H * (W // 128) * (Cin // 8) * 130 * 8, dtype=torch.float16, device=dev)
        out = torch.empty(B, H * W, Cout, device=dev)
        out2 = torch.empty(B, H * W, Cout, device=dev)
        r = torch.randn(B, H * W, Cout, device=dev) if res else None
        stats = torch.zeros(B * Cout * 2, dtype=torch.float64, device=dev)
        bias = torch.zeros(Cout, device=dev)
        front = (x0.data_ptr(), C0, x1.data_ptr() if C1 else 0, C1, s0.data_ptr(), s1.data_ptr() if C1 else 0,
                 gam.data_ptr(), bet.data_ptr(), 0, 0, 8, 1e-6, 1)
        t_gn = timeit(lambda: lib.gn_act_f16(*front, y.data_ptr(), 0, parts, B, H, W, st))
        cands = [(bn, rows) for bn in (128, 64) if Cout % bn == 0 for rows in (4, 2, 1) if H % rows == 0 and rows * bn <= 256]
        line = []
        for bn, rows in cands:
            packed = torch.empty(Cout * Cin * taps * 2, dtype=torch.float16, device=dev)
            ws = 256.0 if parts == 2 else 2.0 ** 16
            lib.pack_conv_weight(w.data_ptr(), packed.data_ptr(), Cout, Cin, taps, bn, rows, parts, ws, st)
            tail = (packed.data_ptr(), bias.data_ptr(), 0 if r is None else r.data_ptr(), 1.0, 1.0 / ws)
            t_c = timeit(lambda: lib.conv_tc(y.data_ptr(), *tail, out2.data_ptr(), stats.data_ptr(), B, H, W, Cin, Cout, taps, 1,
                                             bn, rows, parts, st))
            if taps == 9 and rows == 4:
                line.append(f"bn{bn}R{rows}: conv {t_c:6.1f} fused   n/a")
                continue
            f = lambda: lib.conv_gn_tc(*front, *tail, out.data_ptr(), stats.data_ptr(), B, H, W, Cout, taps, 1, bn, rows, parts, st)
            t_f = timeit(f)
            same = bool(torch.equal(out, out2))
            line.append(f"bn{bn}R{rows}: conv {t_c:6.1f} fused {t_f:6.1f}{'' if same else ' MISMATCH'}")
            for m in masks[1:]:
                lib.conv_set_ablate(m)
                line[-1] += f" m{m}:{timeit(f):6.1f}"
                lib.conv_set_ablate(0)
            for m in (masks if os.environ.get("COUNTERS", "0") == "1" else []):
                dbg = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
                lib.conv_set_ablate(m)
                lib.conv_set_debug(dbg.data_ptr())
                f()
                torch.cuda.synchronize()
                lib.conv_set_debug(0)
                lib.conv_set_ablate(0)
                d = dbg.view(148, 8).double()
                d = d[d[:, 0] > 0].mean(0).tolist()
                print(f"  m{m}", end="")
                print(f"    bn{bn}R{rows} cycles/CTA: mma_total {d[0]:8.0f} waitA {d[1]:7.0f} waitB {d[2]:7.0f} waitAcc {d[3]:7.0f} | "
                      f"epi_total {d[4]:8.0f} epi_wait {d[5]:8.0f} | xf_total {d[7]:8.0f} xf_waitEmptyA {d[6]:8.0f}", flush=True)
        if parts == 3 and taps == 9 and C0 == 64 and C1 == 0 and Cout == 64:      # column walk (rows = 0)
            packed = torch.empty(Cout * Cin * taps * 2, dtype=torch.float16, device=dev)
            ws = 2.0 ** 16
            lib.pack_conv_weight(w.data_ptr(), packed.data_ptr(), Cout, Cin, taps, 64, 0, parts, ws, st)
            tail = (packed.data_ptr(), bias.data_ptr(), 0 if r is None else r.data_ptr(), 1.0, 1.0 / ws)
            f = lambda: lib.conv_gn_tc(*front, *tail, out.data_ptr(), stats.data_ptr(), B, H, W, Cout, taps, 1, 64, 0, parts, st)
            t_f = timeit(f)
            err = float((out - out2).norm() / out2.norm())
            line.append(f"col: fused {t_f:6.1f} (rel vs tile walk {err:.1e})")
            for m in masks[1:]:
                lib.conv_set_ablate(m)
                line[-1] += f" m{m}:{timeit(f):6.1f}"
                lib.conv_set_ablate(0)
            if os.environ.get("COUNTERS", "0") == "1":
                dbg = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
                lib.conv_set_debug(dbg.data_ptr())
                f()
                torch.cuda.synchronize()
                lib.conv_set_debug(0)
                d = dbg.view(148, 8).double()
                d = d[d[:, 0] > 0].mean(0).tolist()
                print(f"    col cycles/CTA: mma_total {d[0]:8.0f} waitHalf {d[1]:7.0f} waitTurn {d[2]:7.0f} waitAcc {d[3]:7.0f} | "
                      f"epi_total {d[4]:8.0f} epi_wait {d[5]:8.0f} | xf_total {d[7]:8.0f} xf_waitEmptyRow {d[6]:8.0f}", flush=True)
        print(f"{H:2d}x{W:<4d} B{B} C{C0}+{C1}->{Cout:<4d} t{taps} res{res} parts{parts}: gn_act {t_gn:6.1f} us | " + " | ".join(line),
              flush=True)


if __name__ == "__main__":
    main()
