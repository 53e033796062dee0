"""Micro-benchmark of the attention entry points on the shapes of the two denoisers (run on the GPU box).
B200_FA_IMPL = tc (default, tcgen05 / TMEM) | mma (mma.sync) | ffma (CUDA cores); COUNTERS=1 prints the in-kernel wait counters."""
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
    ws = torch.empty(lib.flash_attention_workspace(B, heads, T, 0, d, d), dtype=torch.uint8, device=dev)
    ms = timeit(lambda: lib.flash_attention(qkv.data_ptr(), E, out.data_ptr(), W, 2, B, heads, T, 1 / math.sqrt(d), ws.data_ptr(), s))
    fl = 4.0 * B * heads * T * T * d
    print(f"flash_attention    B{B} E{E} h{heads} T{T}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s (algorithmic)")
for B, C, T, W in ((4, 256, 2048, 256), (4, 512, 512, 128)):
    L2 = 13
    heads = C // 32
    qkv = torch.randn(B, T, 3 * C, device=dev)
    pos_p = torch.randn(B, T, C, device=dev)
    kl, pos_l, vl = (torch.randn(B, L2, C, device=dev) for _ in range(3))
    out = torch.empty(2, B * (T // W) * (W // 128) * (C // 8) * 130 * 8, dtype=torch.float16, device=dev)
    ws = torch.empty(lib.flash_attention_workspace(B, heads, T, L2, 64, 32), dtype=torch.uint8, device=dev)
    ms = timeit(lambda: lib.flash_attention_oa(qkv.data_ptr(), pos_p.data_ptr(), kl.data_ptr(), pos_l.data_ptr(),
                                               vl.data_ptr(), out.data_ptr(), W, 2, B, C, heads, T, L2, 1 / math.sqrt(64), ws.data_ptr(), s))
    fl = 2.0 * B * heads * T * (T + L2) * 96
    print(f"flash_attention_oa B{B} C{C} h{heads} T{T}: {ms * 1e3:8.1f} us  {fl / ms / 1e9:7.1f} TFLOP/s (algorithmic)")
    if os.environ.get("COUNTERS") == "1":      # in-kernel cycle counters of the tcgen05 kernel (b200_attn_set_debug), mean over CTAs
        n_cta = (T // 128) * heads * B
        dbg = torch.zeros(n_cta * 16, dtype=torch.int64, device=dev)
        lib.attn_set_debug(dbg.data_ptr())
        lib.flash_attention_oa(qkv.data_ptr(), pos_p.data_ptr(), kl.data_ptr(), pos_l.data_ptr(), vl.data_ptr(), out.data_ptr(), W, 2,
                               B, C, heads, T, L2, 1 / math.sqrt(64), ws.data_ptr(), s)
        torch.cuda.synchronize()
        lib.attn_set_debug(0)
        d = dbg.view(n_cta, 16).double().mean(0).tolist()
        nb = (T + L2 + 127) // 128
        print(f"    per CTA ({nb} key blocks): issuer total {d[0]:.0f} | waits K {d[1]:.0f} S-free {d[2]:.0f} P {d[3]:.0f} V {d[4]:.0f} O-free {d[5]:.0f}")
        print(f"    softmax total {d[6]:.0f} | waits S {d[7]:.0f} O {d[8]:.0f} max-exchange {d[9]:.0f}")
        print(f"    softmax sections: S load {d[13]:.0f} exp {d[14]:.0f} P store {d[15]:.0f} fence+arrive {d[12]:.0f}")
        print(f"    producer total {d[10]:.0f} | waits K-stage {d[11]:.0f}")
print("env B200_FA_IMPL =", os.environ.get("B200_FA_IMPL"))
