#!/usr/bin/env python
"""Small ragged invocations of the CUDA-core kernels added late in round 2, for `compute-sanitizer --tool memcheck`
(no tcgen05 / TMA kernel is launched here; compute-sanitizer is closed on the round-2 GPU pool, so this ran plain): strided local correlation forward + adjoints, the merged lookup adjoint,
gradient ranges and re-zero.      compute-sanitizer --tool memcheck python tools/sanitize_simt.py"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ezflow_b200 as E
from ezflow_b200 import _cabi
from ezflow_b200.similarity.pairwise import grad_layout

dev = torch.device("cuda:0")
torch.manual_seed(0)
E.set_local_corr_mode("exact")  # fp32 CUDA-core kernels only
for (b, c, h, w, p, s) in [(1, 5, 19, 48, 9, 4), (2, 7, 9, 16, 9, 4), (1, 6, 21, 32, 7, 4), (1, 9, 33, 72, 9, 2), (1, 3, 4, 4, 9, 4)]:
    x1 = torch.randn(b, c, h, w, device=dev, requires_grad=True)
    x2 = torch.randn(b, c, h, w, device=dev, requires_grad=True)
    out = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=p, stride=s)(x1, x2)
    out.backward(torch.randn_like(out))
torch.cuda.synchronize()

lib = _cabi.lib()
st = _cabi.stream_ptr(dev)
for (B, H, W, L, r, n) in [(2, 22, 30, 3, 4, 5), (1, 9, 12, 2, 2, 3), (1, 28, 36, 4, 3, 17)]:
    N = H * W
    K = 2 * r + 1
    lay = grad_layout(H, W, L)
    lv = [torch.zeros(B * N * qs, device=dev) for (_, _, _, qs) in lay]
    grid = E.coords_grid(B, H, W, device=dev)
    coords = [(grid + torch.randn(B, 2, H, W, device=dev) * (0.5 + 3 * (k % 3))).contiguous() for k in range(n)]
    coords[-1][:, :, :2] += 500.0
    douts = [torch.randn(B, L * K * K, H, W, device=dev) for _ in range(n)]
    _cabi.check(lib.ezb_corr_lookup_bwd_multi(_cabi.ptr_array(douts), _cabi.ptr_array(coords), n, B, H, W, L, r, _cabi.ptr_array(lv), st))
    cnt = ctypes.c_int64(0)
    _cabi.check(lib.ezb_corr_grad_ranges_count(B, H, W, ctypes.byref(cnt)))
    rg = torch.empty(cnt.value, dtype=torch.int32, device=dev)
    _cabi.check(lib.ezb_corr_grad_ranges(_cabi.ptr_array(coords), n, B, H, W, L, r, _cabi.ptr(rg), st))
    _cabi.check(lib.ezb_corr_grad_rezero(_cabi.ptr_array(lv), _cabi.ptr(rg), B, H, W, L, st))
    torch.cuda.synchronize()
    assert all(float(t.abs().max()) == 0.0 for t in lv), "ranges must cover every window the merged adjoint touched"
print("sanitize_simt ok")
