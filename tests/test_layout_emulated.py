"""Layout-conditioned denoiser (LayoutUnetV1 + LayoutTransformerEncoder + CondContinuousTimeGaussianDiffusion):
oracle pinned to the reference goldens, and the host plan executed through the ABI emulator."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden"))
import lidarcrafter_b200 as L
from abi_emulator import EmulatedLib
from helpers import rel_l2
from lidarcrafter_b200 import _lib
from oracle import unet_torch as O

torch.set_grad_enabled(False)
GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "layout_unet.npz"))

UNET_PARAMS = {'image_size': 32, 'use_fp16': False, 'use_scale_shift_norm': True, 'out_channels': 2, 'model_channels': 64,
               'encoder_channels': 64, 'num_head_channels': 32, 'num_heads': -1, 'num_heads_upsample': -1,
               'num_res_blocks': 2, 'num_attention_blocks': 1, 'resblock_updown': True, 'attention_ds': [4, 8],
               'channel_mult': [1, 2, 4, 8], 'dropout': 0.1, 'use_checkpoint': False,
               'use_positional_embedding_for_attention': True, 'attention_block_type': 'ObjectAwareCrossAttention'}
ENC_PARAMS = {'feature_map_size': [32, 1024], 'used_condition_types': ['obj_class', 'obj_bbox', 'is_valid_obj'],
              'layout_length': 13, 'num_classes_for_layout_object': 9, 'mask_size_for_layout_object': 32, 'hidden_dim': 64,
              'output_dim': 256, 'num_layers': 6, 'num_heads': 4, 'use_final_ln': True, 'use_positional_embedding': False,
              'not_use_layout_fusion_module': False, 'resolution_to_attention': [4, 8], 'use_key_padding_mask': False,
              'out_channels': 10}


def build(cin):
    m = L.unets.__all__["layout_unet_v1"](in_channels=cin, resolution=(32, 1024), **UNET_PARAMS)
    m.coords = L.get_linear_ray_angles(32, 1024, 10, -30)
    sd = O.randomize_state_dict(m.state_dict(), seed=0)
    m.load_state_dict(sd)
    enc = L.unets.__all__["layout_encoder"](**ENC_PARAMS)
    esd = O.randomize_state_dict(enc.state_dict(), seed=1)
    enc.load_state_dict(esd)
    return m.eval(), enc.eval(), sd, esd


def inputs():
    x = torch.randn(1, 2, 32, 1024, generator=torch.Generator().manual_seed(4321))
    return x, torch.tensor([1.3])


@pytest.fixture()
def emu():
    lib = EmulatedLib()
    _lib.set_test_lib(lib)
    yield lib
    _lib.set_test_lib(None)


@pytest.mark.parametrize("name,cin,autoreg", [("layout", 12, False), ("autoreg", 13, True)])
def test_oracle_matches_reference(name, cin, autoreg):
    m, enc, sd, esd = build(cin)
    assert len(sd) == 505 and sum(p.numel() for p in m.parameters()) == (70105602 if cin == 12 else 70106178)
    batch = O.synth_layout_batch(1, seed=0, autoreg=autoreg)
    cond = O.layout_encoder_forward(esd, batch)
    if name == "layout":
        for k in ("xf_proj", "xf_out", "obj_class_embedding", "obj_bbox_embedding"):
            assert rel_l2(cond[k], torch.from_numpy(GOLD["enc_" + k])) < 1e-6
    x, t = inputs()
    y = O.layout_unet_forward(sd, x, t, cond, O.LayoutUnetCfg(in_channels=cin))
    assert rel_l2(y, torch.from_numpy(GOLD[f"{name}_y"])) < 1e-5


@pytest.mark.parametrize("name,cin,autoreg", [("layout", 12, False), ("autoreg", 13, True)])
def test_plan_matches_reference(emu, name, cin, autoreg):
    m, enc, sd, esd = build(cin)
    batch = O.synth_layout_batch(1, seed=0, autoreg=autoreg)
    cond = enc(dict(batch))
    x, t = inputs()
    y = m(x, {"time_condition": t, "other_condition": cond})
    assert rel_l2(y, torch.from_numpy(GOLD[f"{name}_y"])) < 2e-5
    assert emu.calls.count("attention_oa") == 11      # via flash_attention_oa
    assert abs(m.get_plan(1).plan.flops / 1e9 - 255.9) < 6.0       # 255.9 GFLOP / sample-step


def test_cond_sampler_loop(emu):
    m, enc, sd, esd = build(12)
    ddpm = L.CondContinuousTimeGaussianDiffusion(m, enc, prediction_type="eps", noise_schedule="cosine",
                                                 cond_mode="concat")
    assert ddpm.sampling_shape == (2, 32, 1024)
    batch = O.synth_layout_batch(1, seed=3)
    g = torch.Generator().manual_seed(9)
    xs = ddpm.sample(batch, batch_size=1, num_steps=2, progress=False, rng=g, return_all=True, mode="ddim")
    # oracle trajectory with the same noise stream
    g = torch.Generator().manual_seed(9)
    x_T = torch.randn(1, 2, 32, 1024, generator=g)
    cond = O.layout_encoder_forward(esd, batch)
    cfg = O.LayoutUnetCfg(in_channels=12)
    ref = O.sample_uncond(lambda x, l: O.layout_unet_forward(sd, x, l, cond, cfg), x_T, 2, "ddim", 0.0, None,
                          return_all=True)
    assert rel_l2(xs[1], ref[1]) < 2e-4 and rel_l2(xs[-1], ref[-1]) < 1e-3
