"""Golden values of the training LOSS (forward only) from the UNMODIFIED reference (build container only).

    python tests/golden/make_golden_loss.py

ContinuousTimeGaussianDiffusion.forward / p_loss (lidargen/models/diffusion/base.py:119-151, continuous_time.py:135-180)
on the mini EfficientUNet for every objective / criterion; the global torch RNG is seeded right before the call, so a
mirror that draws `torch.rand(B)` (timesteps) and then `randn_like(x_0)` (noise) in the same order reproduces it on the CPU.
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
sys.path.insert(0, HERE)
warnings.filterwarnings("ignore")

CASES = [("eps", "l2", True), ("v", "l1", True), ("x_0", "huber", False)]


def main():
    from make_golden import CASES as UNETS, build_ref_unet
    from oracle import ref_import as R
    torch.set_grad_enabled(False)
    ct = R.continuous_time()
    res, nres, B = UNETS["eunet_mini"]
    m = build_ref_unet(res, nres)
    x0 = torch.randn(B, 2, *res, generator=torch.Generator().manual_seed(99)).clamp(-1, 1)
    mask = (torch.rand(B, 2, *res, generator=torch.Generator().manual_seed(98)) > 0.3).float()
    out = {}
    for obj, crit, minsnr in CASES:
        ddpm = ct.ContinuousTimeGaussianDiffusion(m, prediction_type=obj, loss_type=crit, noise_schedule="cosine",
                                                  min_snr_loss_weight=minsnr)
        torch.manual_seed(4321)
        out[f"{obj}_{crit}_forward"] = np.array([float(ddpm(x0))])
        torch.manual_seed(4321)
        out[f"{obj}_{crit}_forward_masked"] = np.array([float(ddpm(x0, loss_mask=mask))])
        steps = torch.tensor([0.25, 0.8])
        torch.manual_seed(7)
        out[f"{obj}_{crit}_p_loss"] = np.array([float(ddpm.p_loss(x0, steps))])
        out[f"{obj}_{crit}_weight"] = ddpm.get_loss_weight(steps).reshape(-1).numpy()
    np.savez_compressed(os.path.join(HERE, "loss_mini.npz"), **out)
    print({k: v.tolist() for k, v in out.items()})


if __name__ == "__main__":
    main()
