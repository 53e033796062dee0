"""GPU parity of the layout-conditioned denoiser through the public API (registry -> LayoutUnetV1 /
LayoutTransformerEncoder / CondContinuousTimeGaussianDiffusion) against the reference goldens and the oracle."""
import numpy as np
import pytest
import torch

import lidarcrafter_b200 as L
from helpers import rel_l2
from oracle import unet_torch as O
from test_layout_emulated import GOLD, build, inputs

pytestmark = pytest.mark.gpu
torch.set_grad_enabled(False)
TOL = 1e-3


def to_cuda(d):
    return {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in d.items()}


@pytest.mark.parametrize("precision,tol", [("fp16x3", 1e-4), ("fp16f8", 3e-4)])
@pytest.mark.parametrize("name,cin,autoreg", [("layout", 12, False), ("autoreg", 13, True)])
def test_layout_unet_vs_reference_golden(name, cin, autoreg, precision, tol):
    """fp16x3 (the model's default, fp32-grade) and fp16f8 (what `bench.py --workload clip|rollout` runs: fp16 + one e4m3
    correction MMA per conv product; attention keeps the three-term fp16 split in both modes)"""
    m, enc, sd, esd = build(cin)
    m.precision = precision
    m, enc = m.cuda(), enc.cuda()
    batch = to_cuda(O.synth_layout_batch(1, seed=0, autoreg=autoreg))
    cond = enc(dict(batch))
    if name == "layout":
        assert rel_l2(cond["xf_proj"].cpu(), torch.from_numpy(GOLD["enc_xf_proj"])) < 1e-4
    x, t = inputs()
    y = m(x.cuda(), {"time_condition": t.cuda(), "other_condition": cond}).cpu()
    err = rel_l2(y, torch.from_numpy(GOLD[f"{name}_y"]))
    print(name, precision, "rel-L2 vs reference golden:", err)
    assert err < tol


def test_layout_batch4_vs_oracle():
    """config-3 shape (B=4): two samples against the oracle, distinct conditions per sample."""
    m, enc, sd, esd = build(12)
    m, enc = m.cuda(), enc.cuda()
    batch = O.synth_layout_batch(4, seed=7)
    cond = enc(to_cuda(dict(batch)))
    g = torch.Generator().manual_seed(21)
    x = torch.randn(4, 2, 32, 1024, generator=g)
    t = torch.tensor([-5.0, -1.0, 2.0, 8.0])
    y = m(x.cuda(), {"time_condition": t.cuda(), "other_condition": cond}).cpu()
    cond_o = O.layout_encoder_forward(esd, {k: v[2:] for k, v in batch.items()})
    ref = O.layout_unet_forward(sd, x[2:], t[2:], cond_o, O.LayoutUnetCfg(in_channels=12))
    assert rel_l2(y[2:], ref) < TOL


def test_cond_sampler_vs_oracle():
    m, enc, sd, esd = build(12)
    ddpm = L.CondContinuousTimeGaussianDiffusion(m, enc, prediction_type="eps", noise_schedule="cosine",
                                                 cond_mode="concat").cuda()
    batch = O.synth_layout_batch(1, seed=3)
    g = torch.Generator().manual_seed(9)
    x_T = torch.randn(1, 2, 32, 1024, generator=g)
    noises = [torch.randn(1, 2, 32, 1024, generator=g) for _ in range(3)]
    it = iter(noises)
    orig_randn_like = ddpm.randn_like
    ddpm.randn_like = lambda x, rng=None: next(it).to(x.device)
    cond = ddpm.get_network_condition(input_dict=to_cuda(dict(batch)), only_custom_condition=True)
    plan = ddpm.model.get_plan(1)
    plan.set_condition(cond["other_condition"])
    xs = ddpm._sample_from(x_T.cuda(), 3, False, None, True, "ddim", 0.0, plan=plan).cpu()
    cond_o = O.layout_encoder_forward(esd, batch)
    cfg = O.LayoutUnetCfg(in_channels=12)
    ref = O.sample_uncond(lambda x, l: O.layout_unet_forward(sd, x, l, cond_o, cfg), x_T, 3, "ddim", 0.0, None,
                          return_all=True)
    assert rel_l2(xs[1], ref[1]) < TOL and rel_l2(xs[-1], ref[-1]) < TOL
    # public sample(): shapes + the condition is re-folded when the batch dict changes
    ddpm.randn_like = orig_randn_like
    out = ddpm.sample(to_cuda(dict(O.synth_layout_batch(1, seed=4))), batch_size=1, num_steps=2, progress=False,
                      mode="ddim")
    assert out.shape == (1, 2, 32, 1024) and torch.isfinite(out).all()


def test_two_different_layouts_back_to_back_on_one_plan():
    """consecutive sample() calls with DIFFERENT layouts on the same plan -- under inference_mode, with
    the caching allocator handing the second condition the first one's blocks -- must each use their own condition.
    Each one-step sample is checked against the oracle run with that layout."""
    m, enc, sd, esd = build(12)
    ddpm = L.CondContinuousTimeGaussianDiffusion(m, enc, prediction_type="eps", noise_schedule="cosine",
                                                 cond_mode="concat").cuda()
    cfg = O.LayoutUnetCfg(in_channels=12)
    x_T = torch.randn(1, 2, 32, 1024, generator=torch.Generator().manual_seed(9))
    outs = []
    for seed in (3, 4, 3):
        batch = O.synth_layout_batch(1, seed=seed)
        ddpm.randn = lambda *shape, rng=None, **kw: x_T.to(kw.get("device", "cpu"))
        out = ddpm.sample(to_cuda(dict(batch)), batch_size=1, num_steps=1, progress=False, mode="ddim").cpu()
        cond_o = O.layout_encoder_forward(esd, batch)
        ref = O.sample_uncond(lambda x, l: O.layout_unet_forward(sd, x, l, cond_o, cfg), x_T, 1, "ddim", 0.0, None)
        assert rel_l2(out, ref) < TOL, (seed, rel_l2(out, ref))
        outs.append(out)
    assert rel_l2(outs[1], outs[0]) > 1e-2          # the two layouts really give different samples
    assert torch.equal(outs[2], outs[0])


def test_cond_inpaint_vs_oracle_loop():
    """continuous_time_cond.py:283-353 (RePaint with the layout condition) on the GPU: the public inpaint()
    against the same loop written with the oracle's denoiser / DDPM update / q-sample, fed the identical noise stream."""
    m, enc, sd, esd = build(12)
    ddpm = L.CondContinuousTimeGaussianDiffusion(m, enc, prediction_type="eps", noise_schedule="cosine",
                                                 cond_mode="concat").cuda()
    cfg = O.LayoutUnetCfg(in_channels=12)
    batch = O.synth_layout_batch(1, seed=5)
    g = torch.Generator().manual_seed(17)
    known = torch.randn(1, 2, 32, 1024, generator=g).clamp(-1, 1)
    mask = (torch.rand(1, 1, 32, 1024, generator=g) > 0.5).float()
    N = 3
    draws = [torch.randn(1, 2, 32, 1024, generator=g) for _ in range(1 + 2 * N)]
    it = iter(draws)
    ddpm.randn = lambda *shape, rng=None, **kw: next(it).to(kw.get("device", "cpu"))
    ddpm.randn_like = lambda x, rng=None: next(it).to(x.device)
    out = ddpm.inpaint(known.cuda(), mask.cuda(), to_cuda(dict(batch)), num_steps=N, progress=False).cpu()
    # the same loop on the CPU oracle (num_resample_steps = jump_length = 1)
    cond_o = O.layout_encoder_forward(esd, batch)
    it = iter(draws)
    x = next(it)
    steps = torch.linspace(1, 0, N + 1)
    for i in range(N):
        t, s = steps[i:i + 1], steps[i + 1:i + 2]
        lt, ls = O.log_snr_cosine(t), O.log_snr_cosine(s)
        a_s, s_s = O.alpha_sigma(ls)
        known_s = known * a_s.view(-1, 1, 1, 1) + next(it) * s_s.view(-1, 1, 1, 1)
        pred = O.layout_unet_forward(sd, x, lt, cond_o, cfg)
        unknown_s = O.ddpm_update(x, pred, lt, ls, next(it))
        x = mask * known_s + (1 - mask) * unknown_s
    err = rel_l2(out, x)
    print("cond inpaint rel-L2 vs oracle loop:", err)
    assert err < TOL
