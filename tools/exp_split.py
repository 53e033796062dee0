"""Experiment (GPU box): one step graph over the whole batch vs. the batch split into k sub-batches whose plans run on
k streams inside ONE graph (sample trajectories are independent: SURVEY 8e), so that the launch gaps / pipeline
fill / epilogue tails of one chain overlap the other chain's work.

    python tools/exp_split.py [precision] [B]
"""
import os
import sys

import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
from lidarcrafter_b200.efficient_unet import EfficientUNetPlan  # noqa: E402
from lidarcrafter_b200.engine import precision_parts  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "fp16f8"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda")
torch.set_grad_enabled(False)
m, ddpm = bench.build_model(dev, prec)
parts = precision_parts(prec)


def timed_graph(plans, iters=20):
    streams = [torch.cuda.Stream(dev) for _ in plans]

    def body():
        cur = torch.cuda.current_stream(dev)
        if len(plans) == 1:
            plans[0].launch(cur.cuda_stream)
            return
        for p, s in zip(plans, streams):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                p.launch(s.cuda_stream)
        for s in streams:
            cur.wait_stream(s)

    side = torch.cuda.Stream(dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side):
        body()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        body()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


for k in (1, 2, 4):
    if B % k:
        continue
    plans = [EfficientUNetPlan(m, B // k, "tc", parts) for _ in range(k)]
    for p in plans:
        p.x_in.normal_()
        p.t_in.fill_(0.3)
    ms = timed_graph(plans)
    print(f"precision {prec} B {B}: {k} stream(s) x batch {B // k}: {ms:.3f} ms / step  ({B / ms * 1e3:.0f} sample-steps/s)")
    del plans
    torch.cuda.empty_cache()
