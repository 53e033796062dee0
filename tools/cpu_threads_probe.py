"""How many host threads give the best CPU-baseline throughput on this box? (oracle port, B=2 forward)"""
import os, sys, time, json
sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), "..")))
import torch
from oracle import unet_torch as O
import lidarcrafter_b200 as L
torch.set_grad_enabled(False)
RES, NRES = (32, 1024), (3, 3, 3, 3)
m = L.EfficientUNet(in_channels=2, resolution=RES, base_channels=64, channel_multiplier=(1, 2, 4, 8), num_residual_blocks=NRES,
                    gn_num_groups=8, gn_eps=1e-6, attn_num_heads=8, coords_encoding="fourier_features", ring=True)
sd = O.randomize_state_dict(m.state_dict(), 0)
cfg = O.EfficientUNetCfg(resolution=RES, num_residual_blocks=NRES)
x = torch.randn(2, 2, *RES); t = torch.tensor([0.3, 0.3])
out = {"cpu_count": os.cpu_count()}
for n in (8, 16, 32, 64, os.cpu_count()):
    torch.set_num_threads(n)
    O.efficient_unet_forward(sd, x[:1], t[:1], cfg)
    t0 = time.perf_counter(); O.efficient_unet_forward(sd, x, t, cfg); out[str(n)] = time.perf_counter() - t0
print(json.dumps(out))
