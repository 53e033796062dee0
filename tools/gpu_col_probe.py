"""Launch the column-walk fused conv (b200_conv_gn_tc rows = 0) at the bench shape a few times: target of
`ncu --set full --import-source on -k regex:conv_col` (tools/exp_ncu_col.sh)."""
import math
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
from lidarcrafter_b200 import _lib  # noqa: E402


def main():
    lib = _lib.get_lib()
    dev = torch.device("cuda")
    st = torch.cuda.current_stream().cuda_stream
    B, H, W, C = int(os.environ.get("B", "8")), 32, 1024, 64
    res = int(os.environ.get("RES", "1"))
    w = (torch.randn(C, C, 3, 3, device=dev) / math.sqrt(C * 9)).contiguous()
    x0 = torch.randn(B, H * W, C, device=dev)
    s0 = torch.zeros(B, C, 2, dtype=torch.float64, device=dev)
    lib.channel_stats(x0.data_ptr(), s0.data_ptr(), B, H * W, C, st)
    gam, bet = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    out = torch.empty(B, H * W, C, device=dev)
    r = torch.randn(B, H * W, C, device=dev) if res else None
    stats = torch.zeros(B * C * 2, dtype=torch.float64, device=dev)
    bias = torch.zeros(C, device=dev)
    packed = torch.empty(C * C * 9 * 2, dtype=torch.float16, device=dev)
    ws = 2.0 ** 16
    lib.pack_conv_weight(w.data_ptr(), packed.data_ptr(), C, C, 9, 64, 0, 3, ws, st)
    front = (x0.data_ptr(), C, 0, 0, s0.data_ptr(), 0, gam.data_ptr(), bet.data_ptr(), 0, 0, 8, 1e-6, 1)
    for _ in range(int(os.environ.get("N", "4"))):
        lib.conv_gn_tc(*front, packed.data_ptr(), bias.data_ptr(), 0 if r is None else r.data_ptr(), 1.0, 1.0 / ws,
                       out.data_ptr(), stats.data_ptr(), B, H, W, C, 9, 1, 64, 0, 3, st)
    torch.cuda.synchronize()
    print("ok", float(out.abs().mean()))


if __name__ == "__main__":
    main()
