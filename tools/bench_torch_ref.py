#!/usr/bin/env python
"""The honest on-GPU competitor: the reference's own op sequence — plain PyTorch, fp32, TF32 off — restated
from ezflow/similarity (pairwise.py:26-88, resampling.py:56-85, sampler.py:76-129) and timed on the same B200 next to this
repo's kernels.  Test infrastructure / measurement only; nothing here is on the product path.

    python tools/bench_torch_ref.py
"""
import json
import math
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F

import ezflow_b200 as E

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
dev = torch.device("cuda:0")


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters


class TorchPairwise:  # pairwise.py:26-88 + resampling.py:56-85
    def __init__(self, f1, f2, L, r):
        b, c, h, w = f1.shape
        corr = torch.matmul(f1.view(b, c, h * w).transpose(1, 2), f2.view(b, c, h * w)) / math.sqrt(float(c))
        corr = corr.reshape(b * h * w, 1, h, w)
        self.levels = [corr]
        for _ in range(L - 1):
            self.levels.append(F.avg_pool2d(self.levels[-1], 2, stride=2))
        self.r = r

    def __call__(self, coords):
        b, _, h, w = coords.shape
        r = self.r
        c = coords.permute(0, 2, 3, 1).reshape(b * h * w, 1, 1, 2)
        d = torch.linspace(-r, r, 2 * r + 1, device=coords.device)
        delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1).view(1, 2 * r + 1, 2 * r + 1, 2)
        out = []
        for l, lvl in enumerate(self.levels):
            hl, wl = lvl.shape[-2:]
            pos = c / 2**l + delta
            grid = torch.stack([2 * pos[..., 0] / (wl - 1) - 1, 2 * pos[..., 1] / (hl - 1) - 1], dim=-1)
            out.append(F.grid_sample(lvl, grid, align_corners=True).view(b, h, w, -1))
        return torch.cat(out, dim=-1).permute(0, 3, 1, 2).contiguous().float()


def torch_local_corr(x1, x2, P, dp):  # sampler.py:76-129 (stride 1, no padding)
    b, c, h, w = x1.shape
    md = dp * (P - 1) // 2
    x2p = F.pad(x2, (md, md, md, md))
    out = torch.empty((b, P, P, h, w), device=x1.device)
    for i in range(0, 2 * md + 1, dp):
        for j in range(0, 2 * md + 1, dp):
            out[:, i // dp, j // dp] = (x1 * x2p[:, :, i : i + h, j : j + w]).sum(dim=1)
    return out


with torch.no_grad():
    H, W, C = 135, 240, 256
    f1, f2 = torch.randn(1, C, H, W, device=dev), torch.randn(1, C, H, W, device=dev)
    grid = E.coords_grid(1, H, W, device=dev)
    coords = [(grid + torch.randn(1, 2, H, W, device=dev) * (2 + k)).clamp_(-16, W + 16) for k in range(12)]

    def ref_raft():
        fn = TorchPairwise(f1, f2, 4, 4)
        for c in coords:
            o = fn(c)
        return o

    def ours_raft():
        fn = E.MutliScalePairwise4DCorr(f1, f2, num_levels=4, corr_radius=4)
        for c in coords:
            o = fn(c)
        return o

    t_ref, t_new = timed(ref_raft, iters=3, warm=1), timed(ours_raft, iters=10)
    print(json.dumps({"workload": "RAFT 1080p build + 12 lookups, 1 pair", "torch_reference_ops_ms": round(t_ref, 3),
                      "this_repo_ms": round(t_new, 3), "speedup": round(t_ref / t_new, 1)}))
    x1, x2 = torch.randn(8, 256, 48, 64, device=dev), torch.randn(8, 256, 48, 64, device=dev)
    m = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=21, dilation_patch=2)
    t_ref, t_new = timed(lambda: torch_local_corr(x1, x2, 21, 2), iters=3, warm=1), timed(lambda: m(x1, x2), iters=20)
    print(json.dumps({"workload": "cfg2 FlowNetC correlation B8 256x48x64 P21 dp2", "torch_reference_ops_ms": round(t_ref, 3),
                      "this_repo_ms": round(t_new, 3), "speedup": round(t_ref / t_new, 1)}))
    x1, x2 = torch.randn(8, 32, 112, 256, device=dev), torch.randn(8, 32, 112, 256, device=dev)
    m = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=9)
    t_ref, t_new = timed(lambda: torch_local_corr(x1, x2, 9, 1), iters=3, warm=1), timed(lambda: m(x1, x2), iters=20)
    print(json.dumps({"workload": "cfg3 PWC finest level B8 32x112x256 P9", "torch_reference_ops_ms": round(t_ref, 3),
                      "this_repo_ms": round(t_new, 3), "speedup": round(t_ref / t_new, 1)}))
