"""Timing of the layout-conditioned denoiser step (BASELINE.json configs[2]: LayoutUnetV1, batch 4) on one GPU."""
import json, os, sys, time
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..", "tests")))
import torch
import lidarcrafter_b200 as L
from oracle import unet_torch as O
from test_layout_emulated import build
torch.set_grad_enabled(False)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
m, enc, sd, esd = build(12)
ddpm = L.CondContinuousTimeGaussianDiffusion(m, enc, prediction_type="eps", noise_schedule="cosine", cond_mode="concat").cuda()
batch = {k: v.cuda() for k, v in O.synth_layout_batch(B, seed=0).items()}
t0 = time.time(); x = ddpm.sample(batch, batch_size=B, num_steps=3, progress=False, mode="ddim"); torch.cuda.synchronize(); t_first = time.time() - t0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); x = ddpm.sample(batch, batch_size=B, num_steps=20, progress=False, mode="ddim"); e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 20
plan = m.get_plan(B)
prof = plan.plan.profile(torch.cuda.current_stream().cuda_stream, reps=2)
by = {}
for name, ms_k, fl, by_k in prof:
    d = by.setdefault(name, [0.0, 0.0, 0]); d[0] += ms_k; d[1] += fl; d[2] += 1
print(json.dumps({"workload": "LayoutUnetV1 B=%d" % B, "ms_per_step": ms, "sample_steps_per_s": B / ms * 1e3, "first_call_s": t_first,
                  "kernels_per_step": plan.plan.n_kernels, "per_kernel_ms": {k: [round(v[0], 3), v[2]] for k, v in sorted(by.items(), key=lambda kv: -kv[1][0])}}))
