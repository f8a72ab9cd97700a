#!/usr/bin/env python
"""One register-tiled local-correlation call on the finest PWC-Net level (for ncu).  python tools/micro/k4_rt_one.py [C H W]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import ezflow_b200 as E

c, h, w = (int(v) for v in sys.argv[1:4]) if len(sys.argv) >= 4 else (32, 112, 256)
os.environ.setdefault("EZB_LOCAL_CORR_RT", "1")
dev = torch.device("cuda:0")
a, b = torch.randn(8, c, h, w, device=dev), torch.randn(8, c, h, w, device=dev)
m = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=9)
for _ in range(3):
    out = m(a, b)
torch.cuda.synchronize()
print(float(out.abs().mean()))
