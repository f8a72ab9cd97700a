import sys, os
sys.path.insert(0, "/root/repo")
import torch
import ezflow_b200 as E
dev = torch.device("cuda:0")
a = torch.randn(8, 32, 112, 256, device=dev); b = torch.randn(8, 32, 112, 256, device=dev)
m = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=9, padding=0)
for _ in range(3):
    m(a, b)
torch.cuda.synchronize()
