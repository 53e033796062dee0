"""conv_tc vs conv_tc_splitk (K slices on 2 / 4 / 8 CTAs per tile + reduce kernel) for the deep-level shapes at small batch:
us per launch, CUDA events, best of 5 x 10 launches."""
import math
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from lidarcrafter_b200 import _lib  # noqa: E402

SHAPES = [  # B, H, W, Cin, Cout, taps, bn, rows
    (1, 4, 128, 512, 512, 9, 64, 1), (1, 4, 128, 256, 256, 9, 64, 1), (1, 8, 256, 256, 256, 9, 64, 1),
    (1, 8, 256, 128, 128, 9, 64, 1), (1, 8, 256, 512, 128, 9, 64, 1), (1, 4, 128, 512, 1536, 1, 64, 1),
    (4, 4, 128, 512, 512, 9, 128, 1), (8, 4, 128, 256, 256, 9, 64, 1),
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
    lib = _lib.get_lib()
    dev = torch.device("cuda")
    st = torch.cuda.current_stream().cuda_stream
    parts = int(os.environ.get("PARTS", "3"))
    for (B, H, W, Cin, Cout, taps, bn, rows) in SHAPES:
        k = 3 if taps == 9 else 1
        w = (torch.randn(Cout, Cin, k, k, device=dev) / math.sqrt(Cin * taps)).contiguous()
        y = torch.zeros(2, B * H * (W // 128) * (Cin // 8) * 130 * 8, device=dev, dtype=torch.float16)
        y[0] = (torch.randn(y.shape[1], device=dev) * 0.1).half()      # plane 1 (lo / e4m3 pairs) left zero: timing only
        out = torch.empty(B, H * W, Cout, device=dev)
        out2 = torch.empty(B, H * W, Cout, device=dev)
        r = torch.randn(B, H * W, Cout, device=dev)
        stats = torch.zeros(B * Cout * 2, dtype=torch.float64, device=dev)
        bias = torch.zeros(Cout, device=dev)
        packed = torch.empty(Cout * Cin * taps * 2, dtype=torch.float16, device=dev)
        ws_ = 2.0 ** 16 if parts == 3 else 256.0
        lib.pack_conv_weight(w.data_ptr(), packed.data_ptr(), Cout, Cin, taps, bn, rows, parts, ws_, st)
        tail = (packed.data_ptr(), bias.data_ptr(), r.data_ptr(), 1.0, 1.0 / ws_)
        dims = (B, H, W, Cin, Cout, taps, 1, bn, rows, parts, st)
        t0 = timeit(lambda: lib.conv_tc(y.data_ptr(), *tail, out.data_ptr(), stats.data_ptr(), *dims))
        line = f"{H}x{W} B{B} C{Cin}->{Cout} t{taps} bn{bn}R{rows}: conv_tc {t0:6.1f} us"
        for s in (2, 4, 8):
            if (Cin // 16) % s or Cout // 4 > 256 or 256 % (Cout // 4):
                continue
            ws = torch.empty(s, B * H * W * Cout, device=dev)
            t = timeit(lambda: lib.conv_tc_splitk(y.data_ptr(), *tail, out2.data_ptr(), stats.data_ptr(), ws.data_ptr(), s, *dims))
            err = float((out - out2).norm() / out.norm())
            line += f" | x{s}: {t:6.1f} (rel {err:.1e})"
        print(line, flush=True)


if __name__ == "__main__":
    main()
