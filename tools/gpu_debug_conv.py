"""Bring-up diagnostics for the tcgen05 conv kernel (run on the GPU box).  Each case runs in its own
subprocess under a timeout so a trap / hang in one configuration does not hide the others.

    python tools/gpu_debug_conv.py            # all cases
    python tools/gpu_debug_conv.py CASE_JSON  # one case (internal)
"""
import json
import math
import os
import subprocess
import sys

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)

CASES = [
    # taps, Cin, Cout, bn, rows, parts, B, H, W
    dict(taps=1, Cin=32, Cout=64, bn=64, rows=1, parts=3, B=1, H=1, W=128),
    dict(taps=1, Cin=32, Cout=64, bn=64, rows=1, parts=4, B=1, H=1, W=128),
    dict(taps=9, Cin=64, Cout=64, bn=64, rows=2, parts=3, B=1, H=4, W=256),
    dict(taps=9, Cin=64, Cout=64, bn=64, rows=2, parts=4, B=1, H=4, W=256),
    dict(taps=9, Cin=128, Cout=128, bn=128, rows=2, parts=3, B=2, H=4, W=128),
    dict(taps=9, Cin=64, Cout=64, bn=64, rows=2, parts=3, B=8, H=32, W=1024),
    dict(taps=9, Cin=64, Cout=64, bn=64, rows=4, parts=3, B=8, H=32, W=1024),
    dict(taps=9, Cin=64, Cout=64, bn=64, rows=2, parts=4, B=8, H=32, W=1024),
    dict(taps=9, Cin=128, Cout=128, bn=128, rows=2, parts=3, B=8, H=16, W=512),
    dict(taps=9, Cin=128, Cout=128, bn=128, rows=1, parts=4, B=8, H=16, W=512),
    dict(taps=9, Cin=256, Cout=256, bn=128, rows=2, parts=3, B=8, H=8, W=256),
    dict(taps=9, Cin=512, Cout=512, bn=128, rows=1, parts=3, B=8, H=4, W=128),
    dict(taps=9, Cin=512, Cout=512, bn=128, rows=2, parts=3, B=8, H=4, W=128),
    dict(taps=1, Cin=512, Cout=1536, bn=128, rows=2, parts=3, B=8, H=4, W=128),
    dict(taps=9, Cin=64, Cout=64, bn=64, rows=2, parts=2, B=8, H=32, W=1024),
    dict(taps=9, Cin=128, Cout=128, bn=128, rows=2, parts=2, B=8, H=16, W=512),
    dict(taps=9, Cin=256, Cout=256, bn=128, rows=2, parts=2, B=8, H=8, W=256),
    dict(taps=9, Cin=512, Cout=512, bn=128, rows=2, parts=2, B=8, H=4, W=128),
    dict(taps=9, Cin=64, Cout=64, bn=64, rows=4, parts=1, B=8, H=32, W=1024),
    dict(taps=1, Cin=512, Cout=1536, bn=128, rows=2, parts=2, B=8, H=4, W=128),
]


def run_case(c):
    import torch
    import torch.nn.functional as F
    from lidarcrafter_b200 import _lib
    lib = _lib.get_lib()
    _lib.require_b200(0)
    torch.manual_seed(0)
    taps, Cin, Cout, bn, rows, parts, B, H, W = (c[k] for k in ("taps", "Cin", "Cout", "bn", "rows", "parts", "B", "H", "W"))
    k = 3 if taps == 9 else 1
    dev = torch.device("cuda")
    w = (torch.randn(Cout, Cin, k, k, device=dev) / math.sqrt(Cin * taps)).contiguous()
    x = torch.randn(B, H, W, Cin, device=dev).contiguous()
    s = torch.cuda.current_stream().cuda_stream
    parts_op = min(parts, 3)          # parts 4 = parts 3 operands, corrections in their own accumulator columns
    npl = 1 if parts == 1 else 2
    WT = W // 128
    a = torch.zeros(npl, B, H, WT, Cin // 8, 130, 8, dtype=torch.float16, device=dev)   # tile-major operand (+halo pixels)

    def body(t):   # [B, H, WT, G, 130, g] -> [B, H, W, G*g]
        return t[..., 1:129, :].transpose(-3, -2).reshape(B, H, W, -1)
    # operand encoding by the product kernel itself (gn_act without normalisation = cast)
    lib.gn_act_f16(x.data_ptr(), Cin, 0, 0, 0, 0, 0, 0, 0, 0, 1, 0.0, 0, a.data_ptr(), 0, parts_op, B, H, W, s)
    wscale = 2.0 ** ((14 if parts >= 3 else 8) - math.floor(math.log2(float(w.abs().max()))))
    wp = torch.zeros(Cout * Cin * taps * npl, dtype=torch.float16, device=dev)
    out = torch.full((B, H, W, Cout), float("nan"), device=dev)
    st = torch.zeros(B, Cout, 2, dtype=torch.float64, device=dev)
    lib.pack_conv_weight(w.data_ptr(), wp.data_ptr(), Cout, Cin, taps, bn, rows, parts, wscale, s)
    lib.conv_tc(a.data_ptr(), wp.data_ptr(), 0, 0, 1.0, 1.0 / wscale, out.data_ptr(), st.data_ptr(), B, H, W, Cin, Cout,
                taps, 1, bn, rows, parts, s)
    torch.cuda.synchronize()

    def pad(t):
        return F.pad(F.pad(t, (k // 2, k // 2, 0, 0), mode="circular"), (0, 0, k // 2, k // 2)) if k == 3 else t

    def e4m3(v):
        return v.float().clamp(-448, 448).to(torch.float8_e4m3fn).float()

    # reference from the same rounded operands, in fp64 on the GPU
    ws = w * wscale
    whi = ws.half()
    hi = body(a[0]).double()
    if parts <= 2:
        wq = whi.double() + ((ws - whi.float()).half().double() if parts == 2 else 0)
        xq = hi + (body(a[1]).double() if parts == 2 else 0)
        ref = F.conv2d(pad(xq.permute(0, 3, 1, 2)), wq).permute(0, 2, 3, 1) / wscale
    else:
        pair = a[1].contiguous().view(torch.uint8).view(B, H, WT, Cin // 16, 2, 130, 16).view(torch.float8_e4m3fn).double()
        l8, a8 = body(pair[:, :, :, :, 0]), body(pair[:, :, :, :, 1])
        enc = {"hi_plus_l8_vs_x": float(((hi + l8 / 2048) - x.double()).norm() / x.double().norm()),
               "a8_vs_x": float((a8 - x.double()).norm() / x.double().norm())}
        terms = [(hi, whi.double()), (l8, e4m3(ws / 2048).double()), (a8, e4m3(ws - whi.float()).double())]
        ref = sum(F.conv2d(pad(t.permute(0, 3, 1, 2)), wt) for t, wt in terms).permute(0, 2, 3, 1) / wscale
    true = F.conv2d(pad(x.double().permute(0, 3, 1, 2)), w.double()).permute(0, 2, 3, 1)
    err = (out.double() - ref)
    rel = float(err.norm() / ref.norm())
    res = {"case": c, "rel": rel, "rel_vs_fp64_of_fp32_operands": float((out.double() - true).norm() / true.norm()),
           "nan": int(torch.isnan(out).sum()), "maxabs": float(err.abs().nan_to_num(1e9).max())}
    if parts >= 3:
        res["encoding"] = enc
    if not (rel < 1e-4):
        e = err.abs().nan_to_num(1e9)
        res["err_by_px_mod8"] = [float(e[:, :, i::8].mean()) for i in range(8)]
        res["err_by_px_block16"] = [float(e[:, :, i * 16:(i + 1) * 16].mean()) for i in range(min(8, W // 16))]
        res["err_by_ch_mod8"] = [float(e[..., i::8].mean()) for i in range(8)]
        res["err_by_ch_block8"] = [float(e[..., i * 8:(i + 1) * 8].mean()) for i in range(min(8, Cout // 8))]
        res["err_by_row"] = [float(e[:, i].mean()) for i in range(H)]
        res["out_sample"] = out[0, 0, :4, :4].tolist()
        res["ref_sample"] = ref[0, 0, :4, :4].tolist()
        # does the output match a plain (untapped / unshifted) product?  helps spotting descriptor mistakes
        if k == 3:
            center = torch.einsum("bhwk,nk->bhwn", x.double(), w.double()[:, :, 1, 1])
            res["rel_vs_center_tap_only"] = float((out.double() - center).norm() / center.norm())
    ref_st = torch.stack([ref.sum(dim=(1, 2)), (ref ** 2).sum(dim=(1, 2))], -1)
    res["stats_rel"] = float((st - ref_st).norm() / ref_st.norm())
    # timing
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for _ in range(3):
        lib.conv_tc(a.data_ptr(), wp.data_ptr(), 0, 0, 1.0, 1.0 / wscale, out.data_ptr(), 0, B, H, W, Cin, Cout, taps, 1,
                    bn, rows, parts, s)
    e0.record()
    n = 10
    for _ in range(n):
        lib.conv_tc(a.data_ptr(), wp.data_ptr(), 0, 0, 1.0, 1.0 / wscale, out.data_ptr(), 0, B, H, W, Cin, Cout, taps, 1,
                    bn, rows, parts, s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / n
    res["ms"] = ms
    # in-kernel cycle breakdown (one extra launch with the debug counters on)
    dbg = torch.zeros(148 * 8, dtype=torch.int64, device=dev)
    lib.conv_set_debug(dbg.data_ptr())
    lib.conv_tc(a.data_ptr(), wp.data_ptr(), 0, 0, 1.0, 1.0 / wscale, out.data_ptr(), 0, B, H, W, Cin, Cout, taps, 1,
                bn, rows, parts, s)
    torch.cuda.synchronize()
    lib.conv_set_debug(0)
    d = dbg.view(148, 8).double()
    act = d[:, 0] > 0
    if act.any():
        d = d[act]
        names = ["mma_total", "wait_full_a", "wait_full_b", "wait_acc_empty", "epi_total", "epi_wait_acc_full",
                 "prod_wait_empty_a", "prod_wait_empty_b"]
        res["cycles_mean"] = {n_: round(float(d[:, i].mean())) for i, n_ in enumerate(names)}
        res["cycles_max_cta"] = {n_: round(float(d[:, i].max())) for i, n_ in enumerate(names)}
    res["tflops_algorithmic"] = 2.0 * B * H * W * taps * Cin * Cout / ms / 1e9
    print(json.dumps(res))


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1].startswith("{"):
        run_case(json.loads(sys.argv[1]))
        sys.exit(0)
    cases = CASES
    if len(sys.argv) > 1 and sys.argv[1] == "big":      # only the bench-sized shapes
        cases = [c for c in CASES if c["B"] == 8]
    for c in cases:
        try:
            r = subprocess.run([sys.executable, __file__, json.dumps(c)], capture_output=True, text=True, timeout=120)
            tail = (r.stdout.strip().splitlines() or [""])[-1]
            print(tail if tail.startswith("{") else json.dumps({"case": c, "rc": r.returncode,
                                                                 "stdout": r.stdout[-1500:], "stderr": r.stderr[-1500:]}))
        except subprocess.TimeoutExpired:
            print(json.dumps({"case": c, "timeout": True}))
        sys.stdout.flush()
