"""Generate the golden fixtures from the UNMODIFIED reference (run in the build container only).

    python tests/golden/make_golden.py

Imports worldbench/lidarcrafter from /root/reference through oracle/ref_import.py, feeds seeded
inputs and stores inputs-by-seed + outputs as small .npz files next to this script.  Weights are
NOT stored: both sides regenerate them with oracle.unet_torch.randomize_state_dict(seed).
"""
import os
import sys
import warnings

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.abspath(os.path.join(HERE, "..", "..")))
warnings.filterwarnings("ignore")

from oracle import ref_import as R  # noqa: E402
from oracle import unet_torch as O  # noqa: E402

CASES = {
    # name: (resolution, num_residual_blocks, batch)
    "eunet_mini": ((8, 1024), (1, 1, 1, 1), 2),
    "eunet_full": ((32, 1024), (3, 3, 3, 3), 1),
}


def build_ref_unet(resolution, nres):
    eu = R.efficient_unet()
    lid = R.lidar()
    m = eu.EfficientUNet(in_channels=2, resolution=resolution, base_channels=64,
                         channel_multiplier=(1, 2, 4, 8), num_residual_blocks=nres,
                         gn_num_groups=8, gn_eps=1e-6, attn_num_heads=8,
                         coords_encoding="fourier_features", ring=True)
    m.coords = lid.get_linear_ray_angles(resolution[0], resolution[1], 10, -30)
    sd = O.randomize_state_dict(m.state_dict(), seed=0)
    m.load_state_dict(sd)
    return m.eval()


def main():
    torch.manual_seed(0)
    torch.set_grad_enabled(False)
    # ---- UNet forward ----
    for name, (res, nres, B) in CASES.items():
        m = build_ref_unet(res, nres)
        g = torch.Generator().manual_seed(1234)
        x = torch.randn(B, 2, *res, generator=g)
        t = torch.linspace(-6.0, 7.0, B)  # log-SNR values
        y = m(x, t)
        np.savez_compressed(os.path.join(HERE, f"{name}.npz"), y=y.numpy(), t=t.numpy(),
                            x_seed=1234, resolution=np.array(res), nres=np.array(nres))
        print(name, tuple(y.shape), float(y.std()))

    # ---- sampler: 3-step DDIM and 2-step DDPM on the mini model ----
    ct = R.continuous_time()
    m = build_ref_unet(*CASES["eunet_mini"][:2])
    ddpm = ct.ContinuousTimeGaussianDiffusion(m, prediction_type="eps", noise_schedule="cosine")
    out = {}
    for mode, steps, eta in (("ddim", 3, 0.0), ("ddim", 2, 0.5), ("ddpm", 2, 0.0)):
        g = torch.Generator().manual_seed(77)
        xs = ddpm.sample(batch_size=2, num_steps=steps, progress=False, rng=g, return_all=True,
                         mode=mode, ddim_eta=eta)
        out[f"{mode}_{steps}_{eta}_x1"] = xs[1].numpy()      # after the first reverse step
        out[f"{mode}_{steps}_{eta}_last"] = xs[-1].numpy()   # final sample
        print(mode, steps, eta, tuple(xs.shape), float(xs[-1].std()))
    np.savez_compressed(os.path.join(HERE, "sampler_mini.npz"), **out)

    # ---- Resample (FIR) and ring conv building blocks ----
    ops = R.unet_ops()
    g = torch.Generator().manual_seed(5)
    x = torch.randn(2, 3, 6, 16, generator=g)
    np.savez_compressed(os.path.join(HERE, "resample.npz"), x=x.numpy(),
                        down=ops.Resample(down=2, ring=True)(x).numpy(),
                        up=ops.Resample(up=2, ring=True)(x).numpy())

    # ---- schedule ----
    t = torch.linspace(0, 1, 51)
    np.savez_compressed(os.path.join(HERE, "schedule.npz"), t=t.numpy(),
                        log_snr=ct._log_snr_schedule_cosine(t)[:, 0, 0, 0].numpy())
    lid = R.lidar()
    np.savez_compressed(os.path.join(HERE, "ray_angles.npz"),
                        a=lid.get_linear_ray_angles(32, 1024, 10, -30).numpy())


if __name__ == "__main__":
    main()
