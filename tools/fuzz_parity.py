#!/usr/bin/env python
"""Randomised parity sweep (GPU): CUDA path vs the numpy oracle over random shapes / parameters.
    python tools/fuzz_parity.py [n_cases] [seed]
Prints one line per failure and a summary; exit code 1 on any failure."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ezflow_b200 as E
from oracle import similarity_oracle as O

n_cases = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
dev = torch.device("cuda:0")
g = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
fails = 0


def rel(a, b):
    return float(np.linalg.norm(a.astype(np.float64) - b) / max(np.linalg.norm(b), 1e-30)), float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


for i in range(n_cases):
    # ---- all-pairs pyramid + lookup (+ backward on every third case)
    B = int(rng.integers(1, 4)); C = int(rng.choice([1, 3, 17, 64, 100, 256])); L = int(rng.integers(1, 5)); r = int(rng.integers(0, 6))
    H = int(rng.integers(2 << (L - 1), 40)); W = int(rng.integers(2 << (L - 1), 56))
    scale = float(10.0 ** rng.uniform(-3, 3))
    f1 = (rng.standard_normal((B, C, H, W)) * scale).astype(np.float32); f2 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    coords = O.coords_grid(B, H, W) + rng.uniform(-1.5 * r - 2, 1.5 * r + 2, (B, 2, H, W)).astype(np.float32)
    store = "fp16" if rng.random() < 0.7 else "fp32"
    E.set_pyramid_dtype(store)
    ref = O.corr_lookup(O.corr_pyramid(f1, f2, L), coords, r)
    t1, t2 = g(f1).requires_grad_(i % 3 == 0), g(f2).requires_grad_(i % 3 == 0)
    tc = g(coords).requires_grad_(i % 3 == 0)
    out_t = E.MutliScalePairwise4DCorr(t1, t2, num_levels=L, corr_radius=r)(tc)
    rf, mr = rel(out_t.detach().cpu().numpy(), ref)
    ok = rf <= 1e-3 and mr <= 1e-3
    if i % 3 == 0:
        dout = rng.standard_normal(ref.shape).astype(np.float32)
        loss = (out_t * g(dout)).sum()
        dpyr = O.corr_lookup_bwd([(H >> l, W >> l) for l in range(L)], coords, dout, r)
        # 0-3 more lookups through the same pyramid at drifting coordinates (RAFT's iterations): the merged lookup adjoint's
        # shared-memory tiles, its per-lookup fallback, and the support ranges of the pyramid adjoint
        fn_t = out_t.grad_fn  # noqa: F841  (keeps the first lookup's graph alive until backward)
        obj = E.MutliScalePairwise4DCorr(t1, t2, num_levels=L, corr_radius=r)
        loss = (obj(tc) * g(dout)).sum()
        for _ in range(int(rng.integers(0, 4))):
            ck = coords + rng.normal(0, float(rng.choice([0.3, 2.0, 8.0])), coords.shape).astype(np.float32)
            dk = rng.standard_normal(ref.shape).astype(np.float32)
            loss = loss + (obj(g(ck)) * g(dk)).sum()
            dpyr = [a + b for a, b in zip(dpyr, O.corr_lookup_bwd([(H >> l, W >> l) for l in range(L)], ck, dk, r))]
        loss.backward()
        d1, d2 = O.corr_pyramid_bwd(f1, f2, dpyr)
        e1, e2 = rel(t1.grad.cpu().numpy(), d1), rel(t2.grad.cpu().numpy(), d2)
        ec = rel(tc.grad.cpu().numpy(), O.corr_lookup_bwd_coords(O.corr_pyramid(f1, f2, L), coords, dout, r))
        ok = ok and max(e1 + e2 + ec) <= 2e-3
    if not ok:
        fails += 1
        print("FAIL pairwise", (B, C, H, W, L, r, store, scale), rf, mr)
    # ---- local correlation
    B = int(rng.integers(1, 3)); C = int(rng.choice([2, 32, 96, 130, 256])); H = int(rng.integers(3, 30)); W = int(rng.integers(3, 40))
    P = int(rng.choice([1, 3, 5, 9, 21])); s = int(rng.choice([1, 1, 2, 3, 4])); pad = int(rng.choice([0, 0, 1, 4])); dp = int(rng.choice([1, 1, 2, 3, 5]))
    x1 = rng.standard_normal((B, C, H, W)).astype(np.float32); x2 = (rng.standard_normal((B, C, H, W)) * 10.0 ** rng.uniform(-2, 2)).astype(np.float32)
    ref = O.iter_spatial_correlation_sample(x1, x2, patch_size=P, stride=s, padding=pad, dilation_patch=dp)
    for mode, tol in (("tc", 1e-3), ("exact", 2e-5)):
        E.set_local_corr_mode(mode)
        out = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=P, stride=s, padding=pad, dilation_patch=dp)(g(x1), g(x2)).cpu().numpy()
        rf, mr = rel(out, ref) if out.shape == ref.shape else (9, 9)
        if mode == "tc" and int(np.count_nonzero(ref)) < 200:
            tol *= 8  # a handful of random-sign dot products (displacements outside the image are exact zeros and do not count):
            #           error relative to the result is not a stable statistic (tests/test_gpu_local.py keeps one such case with
            #           its bound against the dot-product scale written out)
        if not (rf <= tol and mr <= tol):
            fails += 1
            print("FAIL local", mode, (B, C, H, W, P, s, pad, dp), rf, mr)
    # ---- local-correlation adjoints (K4b): 'tc' = banded tcgen05 GEMMs for the regular cases (1e-3), 'exact' = fp32 kernels (5e-5)
    dout = rng.standard_normal(ref.shape).astype(np.float32)
    d1, d2 = O.iter_spatial_correlation_sample_bwd(x1, x2, dout, patch_size=P, stride=s, padding=pad, dilation_patch=dp)
    for mode, tol in (("tc", 1e-3), ("exact", 5e-5)):
        E.set_local_corr_mode(mode)
        t1, t2 = g(x1).requires_grad_(True), g(x2).requires_grad_(True)
        E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=P, stride=s, padding=pad, dilation_patch=dp)(t1, t2).backward(g(dout))
        e1, e2 = rel(t1.grad.cpu().numpy(), d1), rel(t2.grad.cpu().numpy(), d2)
        if max(e1 + e2) > tol:
            fails += 1
            print("FAIL local bwd", mode, (B, C, H, W, P, s, pad, dp), e1, e2)
    # ---- fused decoder step: warp -> correlation -> LeakyReLU (flows large enough to push taps outside the image)
    flow = (rng.standard_normal((B, 2, H, W)) * rng.choice([0.5, 3.0, 40.0])).astype(np.float32)
    slope = float(rng.choice([0.1, 0.01, 1.0]))
    ref = O.leaky_relu(O.iter_spatial_correlation_sample(x1, O.warp(x2, flow), patch_size=P, stride=s, padding=pad, dilation_patch=dp), slope)
    for mode, tol in (("tc", 1e-3), ("exact", 5e-5)):
        E.set_local_corr_mode(mode)
        m = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=P, stride=s, padding=pad, dilation_patch=dp)
        m.fused_negative_slope = None if slope == 1.0 else slope
        out = m(g(x1), E.lazy_warp(g(x2), g(flow))).cpu().numpy()
        rf, mr = rel(out, ref) if out.shape == ref.shape else (9, 9)
        nnz = int(np.count_nonzero(ref))
        if mode == "tc" and nnz < 200:
            tol *= 8  # as above; pixels the warp masks out are exact zeros and do not average the statistic down
        if nnz < 20 and out.shape == ref.shape:
            # a handful of surviving dot products may all be cancellation-dominated: measure the error against the scale of a
            # C-term dot product of these inputs instead of against the (accidentally small) results
            scale_dp = float(np.sqrt(C) * np.sqrt((x1.astype(np.float64) ** 2).mean()) * np.sqrt((x2.astype(np.float64) ** 2).mean()))
            rf = mr = float(np.abs(out.astype(np.float64) - ref).max() / max(np.abs(ref).max(), scale_dp))
        if np.abs(ref).max() == 0:
            # the warp masks every pixel out: the fused kernel must write exact zeros, not just anything
            rf = mr = 0.0 if (out.shape == ref.shape and not np.any(out)) else 9.0
        if not (rf <= tol and mr <= tol):
            fails += 1
            print("FAIL fused", mode, (B, C, H, W, P, s, pad, dp, slope), rf, mr, "nnz", int(np.count_nonzero(ref)),
                  "max|ref|", float(np.abs(ref).max()), "max|flow|", float(np.abs(flow).max()))
            if os.environ.get("EZB_FUZZ_DUMP"):
                np.savez(os.environ["EZB_FUZZ_DUMP"], x1=x1, x2=x2, flow=flow, ref=ref, out=out, params=np.array([P, s, pad, dp]), slope=slope)
    # ---- convex flow upsampling
    N = int(rng.integers(1, 3)); Cf = int(rng.integers(1, 5)); h = int(rng.integers(1, 12)); w = int(rng.integers(1, 80)); st = int(rng.choice([1, 2, 3, 4, 8, 8]))
    fl = (rng.standard_normal((N, Cf, h, w)) * 10).astype(np.float32); lg = (rng.standard_normal((N, 9 * st * st, h, w)) * 10.0 ** rng.uniform(-1, 1)).astype(np.float32)
    out = E.convex_upsample_flow(g(fl), g(lg), st).cpu().numpy()
    rf, mr = rel(out, O.convex_upsample_flow(fl, lg, st))
    if not (rf <= 2e-5 and mr <= 2e-5):
        fails += 1
        print("FAIL upsample", (N, Cf, h, w, st), rf, mr)
E.set_local_corr_mode("tc"); E.set_pyramid_dtype("fp16")
torch.cuda.synchronize()
print(f"fuzz: {n_cases} cases x (pairwise + local tc/exact + local adjoints + fused warp/activation + convex upsampling), failures: {fails}")
sys.exit(1 if fails else 0)
