"""One FlowNetC-shaped (config 2) local correlation, forward and backward, for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import ezflow_b200 as E
dev = torch.device("cuda:0")
x1 = torch.randn(8, 256, 48, 64, device=dev, requires_grad=True)
x2 = torch.randn(8, 256, 48, 64, device=dev, requires_grad=True)
m = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=21, padding=0, dilation_patch=2)
for _ in range(3):
    out = m(x1, x2)
    out.backward(torch.ones_like(out))
torch.cuda.synchronize()
