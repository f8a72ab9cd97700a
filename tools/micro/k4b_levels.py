#!/usr/bin/env python
"""Adjoint time (fwd+bwd minus fwd) of the local correlation per shape, tensor-core ('tc') vs fp32 CUDA-core ('exact') kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import ezflow_b200 as E

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


cases = [("PWC 196x7x16", (8, 196, 7, 16), dict(patch_size=9)), ("PWC 128x14x32", (8, 128, 14, 32), dict(patch_size=9)),
         ("PWC 96x28x64", (8, 96, 28, 64), dict(patch_size=9)), ("PWC 64x56x128", (8, 64, 56, 128), dict(patch_size=9)),
         ("PWC 32x112x256", (8, 32, 112, 256), dict(patch_size=9)), ("FlowNetC 256x48x64 P21 dp2", (8, 256, 48, 64), dict(patch_size=21, dilation_patch=2)),
         ("DCV s8 d1 256x48x96", (4, 256, 48, 96), dict(patch_size=9)), ("DCV s8 d5", (4, 256, 48, 96), dict(patch_size=9, dilation_patch=5)),
         ("DCV s2 stride 4 128x192x384", (4, 128, 192, 384), dict(patch_size=9, stride=4))]
for name, shape, kw in cases:
    x1 = torch.randn(shape, device=dev, requires_grad=True)
    x2 = torch.randn(shape, device=dev, requires_grad=True)
    m = E.IterSpatialCorrelationSampler(kernel_size=1, padding=0, **kw)
    res = {}
    for mode in ("tc", "exact", "exact_tiled"):
        # 'exact_tiled': the round-1 fp32 tiled kernels (the register-tiled adjoints of local_corr_rt.cu switched off)
        os.environ["EZB_LOCAL_CORR_BWD_RT"] = "0" if mode == "exact_tiled" else "1"
        E.set_local_corr_mode("tc" if mode == "tc" else "exact")
        out = m(x1, x2)
        dout = torch.randn_like(out)
        with torch.no_grad():
            tf = timed(lambda: m(x1.detach(), x2.detach()))

        def fb():
            x1.grad = x2.grad = None
            m(x1, x2).backward(dout)

        res[mode] = (tf, timed(fb) - tf)
    print(f"{name:30s} fwd tc {res['tc'][0]:.3f} exact {res['exact'][0]:.3f} | bwd tc {res['tc'][1]:.3f} exact {res['exact'][1]:.3f} exact_tiled(r1) {res['exact_tiled'][1]:.3f} ms", flush=True)
