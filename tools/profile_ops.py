#!/usr/bin/env python
"""One or two invocations of a named op at its BASELINE size, for ncu captures (never a timing source).

    ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv \\
        --log-file gpurun_out/x.csv python tools/profile_ops.py lookup_conv lookup cfg2 cfg3 cfg4
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ezflow_b200 as E

dev = torch.device("cuda:0")
H, W, C, L, R = 135, 240, 256, 4, 4
P = int(os.environ.get("EZB_PROFILE_PAIRS", "8"))


def raft_inputs():
    g = torch.Generator(device="cpu").manual_seed(1)
    f1 = torch.randn(P, C, H, W, generator=g).to(dev)
    f2 = torch.randn(P, C, H, W, generator=g).to(dev)
    coords = (E.coords_grid(P, H, W) + torch.randn(P, 2, H, W, generator=g) * 4).clamp_(-16, W + 16).to(dev)
    return f1, f2, coords


def lookup_conv():
    f1, f2, coords = raft_inputs()
    w = (torch.randn(256, 324, 1, 1) * 0.05).to(dev)
    b = torch.randn(256).to(dev)
    fn = E.LazyLookupPairwise4DCorr(f1, f2, num_levels=L, corr_radius=R)
    for _ in range(2):
        E.lookup_conv1x1(fn(coords), w, b)


def lookup():
    f1, f2, coords = raft_inputs()
    fn = E.MutliScalePairwise4DCorr(f1, f2, num_levels=L, corr_radius=R)
    for _ in range(2):
        fn(coords)


def cfg2():
    x1, x2 = torch.randn(8, 256, 48, 64, device=dev), torch.randn(8, 256, 48, 64, device=dev)
    m = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=21, padding=0, dilation_patch=2)
    for _ in range(2):
        m(x1, x2)


def cfg3():
    shapes = [(196, 7, 16), (128, 14, 32), (96, 28, 64), (64, 56, 128), (32, 112, 256)]
    m = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=9, padding=0)
    xs = [(torch.randn(8, c, h, w, device=dev), torch.randn(8, c, h, w, device=dev)) for c, h, w in shapes]
    for _ in range(2):
        for a, b in xs:
            m(a, b)


def cfg4():
    xa = [torch.randn(4, 128, 192, 384, device=dev), torch.randn(4, 256, 48, 96, device=dev)]
    xb = [torch.randn(4, 128, 192, 384, device=dev), torch.randn(4, 256, 48, 96, device=dev)]
    m = E.MatryoshkaDilatedCostVolumeList().to(dev)
    for _ in range(2):
        m(xa, xb)


def cfg2_bwd():
    x1 = torch.randn(8, 256, 48, 64, device=dev, requires_grad=True)
    x2 = torch.randn(8, 256, 48, 64, device=dev, requires_grad=True)
    m = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=21, padding=0, dilation_patch=2)
    out = m(x1, x2)
    out.backward(torch.randn_like(out))


def cfg5():
    global P
    P = 1
    f1, f2, coords = raft_inputs()
    f1.requires_grad_(True)
    f2.requires_grad_(True)
    fn = E.MutliScalePairwise4DCorr(f1, f2, num_levels=L, corr_radius=R)
    loss = 0
    for k in range(2):
        loss = loss + (fn(coords) * 0.5).sum()
    loss.backward()


def cfg5_smooth():
    """The bench's smooth-flow training step on 2 pairs: build, 12 lookups, loss, backward (two passes: the second reuses the
    pooled gradient pyramid)."""
    global P
    P = 2
    g = torch.Generator(device="cpu").manual_seed(1)
    f1 = torch.randn(P, C, H, W, generator=g).to(dev).requires_grad_(True)
    f2 = torch.randn(P, C, H, W, generator=g).to(dev).requires_grad_(True)
    grid = E.coords_grid(P, H, W, device=dev)
    flow = torch.nn.functional.interpolate(torch.randn(P, 2, 5, 8, generator=g) * 10, size=(H, W), mode="bicubic", align_corners=True).to(dev)
    coords = [(grid + flow * (1 - 0.5 ** (k + 1)) + torch.randn(P, 2, H, W, generator=g).to(dev) * 0.5).clamp_(-16, W + 16) for k in range(12)]
    douts = [torch.randn(P, L * (2 * R + 1) ** 2, H, W, generator=g).to(dev) for _ in range(12)]
    for _ in range(2):
        f1.grad = f2.grad = None
        fn = E.MutliScalePairwise4DCorr(f1, f2, num_levels=L, corr_radius=R)
        outs = [fn(c) for c in coords]
        torch.autograd.backward(outs, douts)


if __name__ == "__main__":
    with torch.no_grad():
        for name in sys.argv[1:]:
            if name in ("cfg2_bwd", "cfg5", "cfg5_smooth"):
                with torch.enable_grad():
                    globals()[name]()
            else:
                globals()[name]()
    torch.cuda.synchronize()
