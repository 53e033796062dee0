"""Micro-benchmark of the attention entry points on the shapes of the two denoisers (run on the GPU box).
B200_FA_FFMA=1 selects the CUDA-core kernel, default = tensor-core (mma.sync) kernel."""
import math
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from lidarcrafter_b200 import _lib  # noqa: E402

lib = _lib.get_lib()
_lib.require_b200(0)
dev = torch.device("cuda")
s = torch.cuda.current_stream().cuda_stream


def timeit(fn, n=10):
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    e1.synchronize()
    return e0.elapsed_time(e1) / n


for B, E, heads, T, W in ((8, 512, 8, 512, 128), (8, 256, 8, 512, 128)):
    qkv = torch.randn(B, T, 3 * E, device=dev)
    out = torch.empty(2, B * (T // W) * (W // 128) * (E // 8) * 130 * 8, dtype=torch.float16, device=dev)
    d = E // heads
    ms = timeit(lambda: lib.flash_attention(qkv.data_ptr(), E, out.data_ptr(), W, 2, B, heads, T, 1 / math.sqrt(d), s))
    fl = 4.0 * B * heads * T * T * d
    print(f"flash_attention    B{B} E{E} h{heads} T{T}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s (algorithmic)")
for B, C, T, W in ((4, 256, 2048, 256), (4, 512, 512, 128)):
    L2 = 13
    heads = C // 32
    qkv = torch.randn(B, T, 3 * C, device=dev)
    pos_p = torch.randn(B, T, C, device=dev)
    kl, pos_l, vl = (torch.randn(B, L2, C, device=dev) for _ in range(3))
    out = torch.empty(2, B * (T // W) * (W // 128) * (C // 8) * 130 * 8, dtype=torch.float16, device=dev)
    ms = timeit(lambda: lib.flash_attention_oa(qkv.data_ptr(), pos_p.data_ptr(), kl.data_ptr(), pos_l.data_ptr(),
                                               vl.data_ptr(), out.data_ptr(), W, 2, B, C, heads, T, L2, 1 / math.sqrt(64), s))
    fl = 2.0 * B * heads * T * (T + L2) * 96
    print(f"flash_attention_oa B{B} C{C} h{heads} T{T}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s (algorithmic)")
print("env B200_FA_FFMA =", os.environ.get("B200_FA_FFMA"))
