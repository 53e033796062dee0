"""Bottleneck ablations of the tcgen05 conv kernel (diagnostic build: make -C lidarcrafter_b200/csrc EXTRA=-DB200_CONV_ABLATE).
For the dominant shapes, time the kernel with parts of its pipeline switched off (results are wrong by design):
  0 baseline | 1 no FULL_A wait | 2 no FULL_B wait | 3 neither | 4 epilogue without loads/stores | 8 no MMAs issued |
  16 no TMA traffic | combinations.  Prints us per launch (CUDA events, best of 5 x 10 launches)."""
import math
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from lidarcrafter_b200 import _lib  # noqa: E402

SHAPES = [  # B, H, W, Cin, Cout, bn, rows, res, taps
    (8, 32, 1024, 64, 64, 64, 2, 0, 9),
    (8, 32, 1024, 64, 64, 64, 2, 1, 9),
    (8, 32, 1024, 64, 64, 64, 1, 0, 9),
    (8, 16, 512, 64, 64, 64, 2, 0, 9),
    (8, 16, 512, 128, 128, 128, 1, 0, 9),
    (8, 8, 256, 256, 256, 128, 1, 0, 9),
    (8, 4, 128, 512, 512, 128, 1, 0, 9),
    (8, 4, 128, 256, 256, 64, 1, 0, 9),
    (8, 4, 128, 256, 256, 128, 1, 0, 9),
    (8, 8, 256, 128, 128, 128, 1, 0, 9),
    (8, 8, 256, 128, 128, 64, 1, 0, 9),
    (8, 8, 256, 128, 128, 64, 2, 0, 9),
]
MASKS = [0, 16]


def main():
    lib = _lib.get_lib()
    _lib.require_b200(0)
    dev = torch.device("cuda")
    st = torch.cuda.current_stream().cuda_stream
    parts = 2
    for (B, H, W, Cin, Cout, bn, rows, res, taps) in SHAPES:
        kk = 3 if taps == 9 else 1
        w = (torch.randn(Cout, Cin, kk, kk, device=dev) / math.sqrt(Cin * taps)).contiguous()
        packed = torch.empty(Cout * Cin * taps * 2, dtype=torch.float16, device=dev)
        lib.pack_conv_weight(w.data_ptr(), packed.data_ptr(), Cout, Cin, taps, bn, rows, parts, 256.0, st)
        a = (torch.randn(2, B * H * (W // 128) * (Cin // 8) * 130 * 8, device=dev) * 0.5).half()
        out = torch.empty(B, H * W, Cout, device=dev)
        r = torch.randn(B, H * W, Cout, device=dev) if res else None
        stats = torch.zeros(B * Cout * 2, dtype=torch.float64, device=dev)
        bias = torch.zeros(Cout, device=dev)
        args = (a.data_ptr(), packed.data_ptr(), bias.data_ptr(), 0 if r is None else r.data_ptr(), 1.0, 1.0 / 256.0,
                out.data_ptr(), stats.data_ptr(), B, H, W, Cin, Cout, taps, 1, bn, rows, parts, st)
        line = []
        for m in MASKS:
            lib.conv_set_ablate(m)
            lib.conv_tc(*args)
            torch.cuda.synchronize()
            best = 1e9
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    lib.conv_tc(*args)
                e1.record()
                e1.synchronize()
                best = min(best, e0.elapsed_time(e1) / 10 * 1e3)
            line.append(f"m{m}:{best:6.1f}")
            if os.environ.get("COUNTERS", "0") == "1":
                dbg = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
                lib.conv_set_debug(dbg.data_ptr())
                lib.conv_tc(*args)
                torch.cuda.synchronize()
                lib.conv_set_debug(0)
                d = dbg.view(148, 8).double()
                d = d[d[:, 0] > 0].mean(0).tolist()
                print(f"    m{m:<2d} cycles/CTA: mma_total {d[0]:8.0f} waitA {d[1]:7.0f} waitB {d[2]:7.0f} waitAcc {d[3]:7.0f} | "
                      f"epi_total {d[4]:8.0f} epi_wait {d[5]:8.0f} | prod waitEmptyA {d[6]:8.0f} waitEmptyB {d[7]:8.0f}", flush=True)
        lib.conv_set_ablate(0)
        fl = 2.0 * B * H * W * taps * Cin * Cout
        print(f"{H:2d}x{W:<4d} C{Cin:<3d}->{Cout:<4d} t{taps} bn{bn} R{rows} res{res} ({fl / 1e9:5.1f} GF) us:", "  ".join(line), flush=True)


if __name__ == "__main__":
    main()
