"""DCVNet's coarse cost-volume scale alone ((4,256,48,96), dilations 1,2,3,5,9,16, 9x9 window), three calls: for ncu captures."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import ezflow_b200 as E

dev = torch.device("cuda:0")
x1, x2 = torch.randn(4, 256, 48, 96, device=dev), torch.randn(4, 256, 48, 96, device=dev)
m = E.MatryoshkaDilatedCostVolume(num_groups=1, max_displacement=4, stride=1, dilations=[1, 2, 3, 5, 9, 16]).to(dev)
with torch.no_grad():
    for _ in range(3):
        m(x1, x2)
torch.cuda.synchronize()
