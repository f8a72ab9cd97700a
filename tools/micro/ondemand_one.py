"""One on-demand lookup of 2 1080p pairs on a smooth flow (three calls): for ncu captures of corr_ondemand_tiled_kernel."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import ezflow_b200 as E

dev = torch.device("cuda:0")
g = torch.Generator(device="cpu").manual_seed(5)
H, W, C = 135, 240, 256
f1, f2 = torch.randn(2, C, H, W, generator=g).to(dev), torch.randn(2, C, H, W, generator=g).to(dev)
flow = torch.nn.functional.interpolate(torch.randn(2, 2, 5, 8, generator=g) * 10, size=(H, W), mode="bicubic", align_corners=True)
coords = (E.coords_grid(2, H, W) + flow + torch.randn(2, 2, H, W, generator=g) * 0.5).to(dev)
with torch.no_grad():
    od = E.OnDemandPairwise4DCorr(f1, f2, num_levels=4, corr_radius=4)
    for _ in range(3):
        od(coords)
torch.cuda.synchronize()
