"""Micro-benchmark of b200_gn_act_f16 on the shapes of the EfficientUNet step (run on the GPU box)."""
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from lidarcrafter_b200 import _lib  # noqa: E402

lib = _lib.get_lib()
_lib.require_b200(0)
dev = torch.device("cuda")
B = 8
SHAPES = [(32, 1024, 64, 0), (32, 1024, 128, 1), (16, 512, 128, 0), (8, 256, 256, 0), (4, 128, 512, 0)]
parts = int(sys.argv[1]) if len(sys.argv) > 1 else 3
silu = int(sys.argv[2]) if len(sys.argv) > 2 else 1
tot = 0.0
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for H, W, C, raw in SHAPES:
    x = torch.randn(B, H * W, C, device=dev)
    st = torch.stack([x.double().sum(1), (x.double() ** 2).sum(1)], -1).contiguous()
    gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    n = B * H * (W // 128) * (C // 8) * 130 * 8
    y = torch.empty(2, n, dtype=torch.float16, device=dev)
    yr = torch.empty(2, n, dtype=torch.float16, device=dev)
    s = torch.cuda.current_stream().cuda_stream
    args = (x.data_ptr(), C, 0, 0, st.data_ptr(), 0, gam.data_ptr(), bet.data_ptr(), 0, 0, 8, 1e-6, silu, y.data_ptr(),
            yr.data_ptr() if raw else 0, parts, B, H, W, s)
    for _ in range(3):
        lib.gn_act_f16(*args)
    ts = []
    for _ in range(20):
        flush.zero_()                                   # evict L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        lib.gn_act_f16(*args)
        e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    by = B * H * W * C * (4 + 4 * (2 if raw else 1))
    tot += ms
    print(f"gn_act {H}x{W} C{C} raw{raw}: {ms * 1e3:7.1f} us  {by / ms / 1e6:7.0f} GB/s")
print(f"parts {parts} silu {silu} sum {tot * 1e3:.1f} us  (env V1={os.environ.get('B200_GN_V1')} BPS={os.environ.get('B200_GN_BPS')})")
