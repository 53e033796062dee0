"""Shared test utilities (model construction with deterministic weights, golden loading)."""
import os

import numpy as np
import torch

import lidarcrafter_b200 as L
from oracle import unet_torch as O

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = {"eunet_mini": ((8, 1024), (1, 1, 1, 1), 2), "eunet_full": ((32, 1024), (3, 3, 3, 3), 1)}


def make_unet(res, nres, seed=0):
    m = L.EfficientUNet(in_channels=2, resolution=res, base_channels=64, channel_multiplier=(1, 2, 4, 8),
                        num_residual_blocks=nres, gn_num_groups=8, gn_eps=1e-6, attn_num_heads=8,
                        coords_encoding="fourier_features", ring=True)
    m.coords = L.get_linear_ray_angles(res[0], res[1], 10, -30)
    sd = O.randomize_state_dict(m.state_dict(), seed=seed)
    m.load_state_dict(sd)
    return m.eval(), sd


def golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def golden_inputs(name):
    res, nres, B = CASES[name]
    d = golden(name)
    g = torch.Generator().manual_seed(int(d["x_seed"]))
    x = torch.randn(B, 2, *res, generator=g)
    return x, torch.from_numpy(d["t"]), torch.from_numpy(d["y"])


def rel_l2(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())
