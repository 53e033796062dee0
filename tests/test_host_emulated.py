"""Host logic without a GPU: the static kernel plans (layer wiring, weight packing order, AdaGN offsets,
statistics slots, sampler loop) executed through tests/abi_emulator.py and compared with the golden
outputs of the unmodified reference."""
import pytest
import torch

from abi_emulator import EmulatedLib
from helpers import CASES, golden, golden_inputs, rel_l2
from helpers import make_unet as _make_unet
from lidarcrafter_b200 import _lib
import lidarcrafter_b200 as L

torch.set_grad_enabled(False)


def make_unet(res, nres, precision="fp16x3"):
    """host-logic tests run the fp32-grade mode (their tolerances check wiring, not the fp16f8 default's 5e-5)"""
    m, sd = _make_unet(res, nres)
    m.precision = precision
    return m, sd


@pytest.fixture()
def emu():
    lib = EmulatedLib()
    _lib.set_test_lib(lib)
    yield lib
    _lib.set_test_lib(None)


@pytest.mark.parametrize("precision,tol", [("fp16x3", 2e-5), ("fp16f8", 2e-4), ("fp16", 4e-3)])
def test_plan_mini_matches_reference(emu, precision, tol):
    res, nres, B = CASES["eunet_mini"]
    m, _ = make_unet(res, nres)
    m.precision = precision
    x, t, y_ref = golden_inputs("eunet_mini")
    assert rel_l2(m(x, t), y_ref) < tol
    if precision == "fp16":        # single-pass fp16 (measurement only) keeps the separate gn_act launches
        assert emu.calls.count("conv_tc") == 30
    else:                          # 25 convs with the GroupNorm / cast front end fused in, 5 on pre-built operands
        assert emu.calls.count("conv_gn_tc") + emu.calls.count("conv_tc") == 30 and emu.calls.count("gn_act_f16") == 0


def test_plan_mini_unfused_front_matches(emu, monkeypatch):
    """B200_FUSE_FRONT=0: separate gn_act launches + conv_tc instead of the fused b200_conv_gn_tc, same result"""
    monkeypatch.setenv("B200_FUSE_FRONT", "0")
    res, nres, B = CASES["eunet_mini"]
    m, _ = make_unet(res, nres)
    x, t, y_ref = golden_inputs("eunet_mini")
    assert rel_l2(m(x, t), y_ref) < 2e-5
    assert emu.calls.count("conv_gn_tc") == 0 and emu.calls.count("conv_tc") == 30 and emu.calls.count("gn_act_f16") > 16


def test_plan_full_matches_reference(emu):
    res, nres, B = CASES["eunet_full"]
    m, _ = make_unet(res, nres)
    x, t, y_ref = golden_inputs("eunet_full")
    assert rel_l2(m(x, t), y_ref) < 2e-5
    plan = m.get_plan(1)
    assert abs(plan.plan.flops / 1e9 - 116.6) < 3.0     # 116.6 GFLOP / sample-step


def test_ffma_crosscheck_plan(emu):
    res, nres, B = CASES["eunet_mini"]
    m, _ = make_unet(res, nres)
    m.conv_impl = "ffma"
    x, t, y_ref = golden_inputs("eunet_mini")
    assert rel_l2(m(x, t), y_ref) < 2e-5
    assert emu.calls.count("conv_ffma") == 30


def test_state_dict_reload_invalidates_plans(emu):
    res, nres, B = CASES["eunet_mini"]
    m, sd = make_unet(res, nres)
    x, t, y_ref = golden_inputs("eunet_mini")
    y0 = m(x, t)
    sd2 = {k: (v * 0.5 if k == "out_conv.weight" else v) for k, v in sd.items()}
    m.load_state_dict(sd2)
    y1 = m(x, t)
    assert rel_l2(y1, y0) > 1e-2


@pytest.mark.parametrize("mode,steps,eta", [("ddim", 3, 0.0), ("ddim", 2, 0.5), ("ddpm", 2, 0.0)])
def test_sampler_matches_reference(emu, mode, steps, eta):
    res, nres, _ = CASES["eunet_mini"]
    m, _ = make_unet(res, nres)
    ddpm = L.ContinuousTimeGaussianDiffusion(m, prediction_type="eps", noise_schedule="cosine")
    g = torch.Generator().manual_seed(77)
    xs = ddpm.sample(batch_size=2, num_steps=steps, progress=False, rng=g, return_all=True, mode=mode, ddim_eta=eta)
    d = golden("sampler_mini")
    assert rel_l2(xs[1], torch.from_numpy(d[f"{mode}_{steps}_{eta}_x1"])) < 2e-4
    assert rel_l2(xs[-1], torch.from_numpy(d[f"{mode}_{steps}_{eta}_last"])) < 1e-3
    # alias required by BASELINE.json's wording
    assert type(ddpm).p_sample_loop is type(ddpm).sample


def test_p_step_signature_and_result(emu):
    from oracle import unet_torch as O
    res, nres, _ = CASES["eunet_mini"]
    m, sd = make_unet(res, nres)
    ddpm = L.ContinuousTimeGaussianDiffusion(m, prediction_type="eps", noise_schedule="cosine")
    g = torch.Generator().manual_seed(3)
    x = torch.randn(2, 2, *res, generator=g)
    st, ss = torch.tensor([0.7, 0.4]), torch.tensor([0.6, 0.3])
    y = ddpm.p_step(x, st, ss, rng=torch.Generator().manual_seed(5), mode="ddim", ddim_eta=0.0)
    cfg = O.EfficientUNetCfg(resolution=res, num_residual_blocks=nres)
    lt, ls = O.log_snr_cosine(st), O.log_snr_cosine(ss)
    ref = O.ddim_update(x, O.efficient_unet_forward(sd, x, lt, cfg), lt, ls)
    assert rel_l2(y, ref) < 2e-5


def test_registry_keys():
    assert L.unets.__all__["efficient_unet"] is L.EfficientUNet


def test_column_walk_plan_matches_tile_walk(emu, monkeypatch):
    """a plan whose tuned tile for the 64 -> 64 layers is (bn 64, rows 0) -- the column-walk schedule of b200_conv_gn_tc with its
    own stacked-filter-row weight image (csrc/conv_col.cuh) -- gives the forward of the tile-walk plan, and the eligibility
    rule mirrors the kernel's argument checks"""
    from lidarcrafter_b200 import engine
    assert engine.col_walk_ok(8, 32, 1024, 64, 0, 64, 9, 3) and engine.col_walk_ok(1, 32, 1024, 64, 0, 64, 9, 3)
    assert not engine.col_walk_ok(8, 32, 1024, 64, 0, 64, 9, 2)          # fp16x3: the weights do not fit
    assert not engine.col_walk_ok(8, 32, 1024, 64, 64, 64, 9, 3)         # channel concat
    assert not engine.col_walk_ok(8, 32, 1024, 64, 0, 128, 9, 3) and not engine.col_walk_ok(8, 32, 1024, 64, 0, 64, 1, 3)
    assert not engine.col_walk_ok(400, 4, 128, 64, 0, 64, 9, 3)          # a CTA's run would span more than two samples
    res, nres, B = CASES["eunet_mini"]
    x, t, y_ref = golden_inputs("eunet_mini")
    m, _ = make_unet(res, nres, precision="fp16f8")
    y0 = m(x, t)
    saved = dict(engine._TUNE_CACHE)
    try:
        engine._TUNE_CACHE.clear()
        H, W = res
        for has_res in (False, True):
            engine._TUNE_CACHE[(B, H, W, 64, 64, 9, 3, has_res, True)] = (64, 0)
        m2, _ = make_unet(res, nres, precision="fp16f8")
        emu.calls.clear()
        y1 = m2(x, t)
        rows0 = [a for fn, a in m2.get_plan(B).plan.ops if getattr(fn, "__name__", "") == "conv_gn_tc" and a[27] == 0]
        assert len(rows0) >= 2, "no layer of the plan runs the column walk"
    finally:
        engine._TUNE_CACHE.clear()
        engine._TUNE_CACHE.update(saved)
    assert rel_l2(y1, y0) < 1e-6 and rel_l2(y1, y_ref) < 2e-4


def test_split_k_plan_matches_plain_plan(emu):
    """a plan whose tuner chose split-K for the deepest 3x3 convs (b200_conv_tc_splitk + workspace) gives the forward of the
    plain plan"""
    from lidarcrafter_b200 import engine
    res, nres, B = CASES["eunet_mini"]
    x, t, y_ref = golden_inputs("eunet_mini")
    m, _ = make_unet(res, nres)
    y0 = m(x, t)
    shapes = {(a[9], a[10], a[11], a[12], a[13], a[15], a[16], a[3] != 0) for fn, a in m.get_plan(B).plan.ops
              if getattr(fn, "__name__", "") == "conv_tc" and a[11] >= 128 and a[13] == 9}
    assert shapes
    saved = dict(engine._TUNE_CACHE)
    try:
        for (H, W, Cin, Cout, taps, bn, rows, has_res) in shapes:
            engine._TUNE_CACHE[("split", B, H, W, Cin, Cout, taps, 2, has_res, bn, rows)] = (2, 0)
        m2, _ = make_unet(res, nres)
        emu.calls.clear()
        y1 = m2(x, t)
        assert emu.calls.count("conv_tc_splitk") >= len(shapes)
    finally:
        engine._TUNE_CACHE.clear()
        engine._TUNE_CACHE.update(saved)
    assert rel_l2(y1, y0) < 1e-6 and rel_l2(y1, y_ref) < 2e-5


def test_tile_picker_respects_kernel_limits():
    from lidarcrafter_b200.engine import pick_tile
    for B in (1, 2, 8, 64):
        for (H, W, C) in ((32, 1024, 64), (16, 512, 128), (8, 256, 256), (4, 128, 512), (1, 128, 1536)):
            for taps in (1, 9):
                for parts in (1, 2):
                    bn, rows = pick_tile(B, H, W, C, taps, parts)
                    assert C % bn == 0 and H % rows == 0
                    assert rows * bn <= 256


@pytest.mark.parametrize("nres,jump", [(1, 1), (2, 2)])
def test_repaint_matches_reference(emu, nres, jump):
    """RePaint loop (continuous_time.py:262-319) against the reference run stored in tests/golden/repaint_mini.npz
    (generated with the same CPU generator, so the whole noise stream is identical)."""
    res, nres_blocks, _ = CASES["eunet_mini"]
    m, _ = make_unet(res, nres_blocks)
    ddpm = L.ContinuousTimeGaussianDiffusion(m, prediction_type="eps", noise_schedule="cosine")
    g = torch.Generator().manual_seed(31)
    known = torch.rand(2, 2, 8, 1024, generator=g) * 2 - 1
    mask = (torch.rand(2, 1, 8, 1024, generator=g) > 0.5).float()
    x = ddpm.repaint(known, mask, num_steps=2, num_resample_steps=nres, jump_length=jump, progress=False,
                     rng=torch.Generator().manual_seed(55))
    ref = torch.from_numpy(golden("repaint_mini")[f"repaint_{nres}_{jump}"])
    assert rel_l2(x, ref) < 1e-3


def test_tile_choices_persist_through_the_tune_file(tmp_path, monkeypatch):
    """B200_TUNE_FILE: the measured (bn, rows) choices survive a process boundary (a profiled run must execute the
    tiles the plain run measured); keys are (B, H, W, Cin, Cout, taps, parts, has_residual)."""
    from lidarcrafter_b200 import engine
    path = tmp_path / "tiles.json"
    monkeypatch.setenv("B200_TUNE_FILE", str(path))
    saved = dict(engine._TUNE_CACHE)
    try:
        engine._TUNE_CACHE.clear()
        engine._TUNE_CACHE[(8, 32, 1024, 64, 64, 9, 2, True)] = (64, 2)
        engine._TUNE_CACHE[(8, 4, 128, 512, 512, 9, 2, False)] = (128, 1)
        engine._save_tune_file()
        engine._TUNE_CACHE.clear()
        engine._TUNE_FILE_LOADED = False
        engine._load_tune_file()
        assert engine._TUNE_CACHE == {(8, 32, 1024, 64, 64, 9, 2, True): (64, 2), (8, 4, 128, 512, 512, 9, 2, False): (128, 1)}
    finally:
        engine._TUNE_CACHE.clear()
        engine._TUNE_CACHE.update(saved)
        engine._TUNE_FILE_LOADED = False


@pytest.mark.parametrize("objective,criterion,minsnr", [("eps", "l2", True), ("v", "l1", True), ("x_0", "huber", False)])
def test_loss_evaluation_matches_reference(emu, objective, criterion, minsnr):
    """forward() / p_loss() (no gradient) against the reference's loss values (tests/golden/loss_mini.npz): same RNG draw
    order (timesteps, then noise), same min-SNR weighting and -- deliberately -- the reference's [B,1] x [B,1,1,1] broadcast
    of loss and weight (base.py:150)."""
    res, nres, B = CASES["eunet_mini"]
    m, _ = make_unet(res, nres)
    d = golden("loss_mini")
    ddpm = L.ContinuousTimeGaussianDiffusion(m, prediction_type=objective, loss_type=criterion, noise_schedule="cosine",
                                             min_snr_loss_weight=minsnr)
    x0 = torch.randn(B, 2, *res, generator=torch.Generator().manual_seed(99)).clamp(-1, 1)
    mask = (torch.rand(B, 2, *res, generator=torch.Generator().manual_seed(98)) > 0.3).float()
    key = f"{objective}_{criterion}"
    torch.manual_seed(4321)
    assert abs(float(ddpm(x0)) - d[key + "_forward"][0]) < 2e-4 * abs(d[key + "_forward"][0])
    torch.manual_seed(4321)
    assert abs(float(ddpm(x0, loss_mask=mask)) - d[key + "_forward_masked"][0]) < 2e-4 * abs(d[key + "_forward_masked"][0])
    steps = torch.tensor([0.25, 0.8])
    torch.manual_seed(7)
    assert abs(float(ddpm.p_loss(x0, steps)) - d[key + "_p_loss"][0]) < 2e-4 * abs(d[key + "_p_loss"][0])
    assert torch.allclose(ddpm.get_loss_weight(steps).reshape(-1), torch.from_numpy(d[key + "_weight"]), rtol=1e-5)
