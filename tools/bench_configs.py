#!/usr/bin/env python
"""Times the BASELINE configs 2-4 (FlowNetC / PWC-Net / DCVNet similarity ops) and the RAFT training path
(config 5: lookup backward + pyramid backward for one pair) on one GPU with CUDA events; prints achieved GB/s
against the algorithmic bytes of BASELINE.md §3.  Parity for the same shapes is in tests/test_gpu_local.py.

    python tools/bench_configs.py [cfg2 cfg3 cfg4 cfg5 upsample bwd]
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ezflow_b200 as E

dev = torch.device("cuda:0")
PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
    os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6556.8
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()  # > L2 between timed iterations
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def report(name, ms, nbytes, flops):
    print(json.dumps({"config": name, "ms": round(ms, 4), "algorithmic_MB": round(nbytes / 1e6, 1), "GBps": round(nbytes / ms / 1e6, 1),
                      "frac_of_hbm_peak": round(nbytes / ms / 1e6 / PEAK, 3), "TFLOPs": round(flops / ms / 1e9, 2)}), flush=True)


def cfg2():
    x1, x2 = torch.randn(8, 256, 48, 64, device=dev), torch.randn(8, 256, 48, 64, device=dev)
    m = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=21, padding=0, dilation_patch=2)
    with torch.no_grad():
        ms = timed(lambda: m(x1, x2))
    report("cfg2 FlowNetC corr B8 256x48x64 P21 dp2", ms, 2 * x1.numel() * 4 + 8 * 441 * 48 * 64 * 4, 2.0 * 8 * 441 * 48 * 64 * 256)


def cfg3():
    shapes = [(196, 7, 16), (128, 14, 32), (96, 28, 64), (64, 56, 128), (32, 112, 256)]
    xs = [(torch.randn(8, c, h, w, device=dev), torch.randn(8, c, h, w, device=dev), torch.randn(8, 2, h, w, device=dev) * 2) for c, h, w in shapes]
    m = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=9, padding=0)

    def run(with_warp):
        for i, (a, b, f) in enumerate(xs):
            bb = E.warp(b, f) if (with_warp and i > 0) else b
            m(a, bb)

    nb = sum(2 * a.numel() * 4 + 8 * 81 * a.shape[2] * a.shape[3] * 4 for a, _, _ in xs)
    fl = sum(2.0 * 8 * 81 * a.shape[2] * a.shape[3] * a.shape[1] for a, _, _ in xs)
    with torch.no_grad():
        report("cfg3 PWC 5-level corr B8 (no warp)", timed(lambda: run(False)), nb, fl)
        nbw = nb + sum(2 * a.numel() * 4 + f.numel() * 4 for a, _, f in xs[1:])
        report("cfg3 PWC 5-level warp+corr B8", timed(lambda: run(True)), nbw, fl)

        # the decoder step as the reference runs it (pyramid.py:98-102,138-141): warp -> corr -> view -> LeakyReLU(0.1)
        def unfused():
            for i, (a, b, f) in enumerate(xs):
                c = m(a, E.warp(b, f) if i > 0 else b)
                torch.nn.functional.leaky_relu(c.view(c.shape[0], -1, c.shape[3], c.shape[4]), 0.1)

        mf = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=9, padding=0)
        mf.fused_negative_slope = 0.1

        def fused():
            for i, (a, b, f) in enumerate(xs):
                c = mf(a, E.lazy_warp(b, f) if i > 0 else b)
                c.view(c.shape[0], -1, c.shape[3], c.shape[4])

        nbf = nb + sum(f.numel() * 4 for _, _, f in xs[1:])  # fused: the warped tensor is never written or re-read
        report("cfg3 PWC warp+corr+LeakyReLU B8, separate kernels", timed(unfused), nbw + 2 * (nb - sum(2 * a.numel() * 4 for a, _, _ in xs)), fl)
        report("cfg3 PWC warp+corr+LeakyReLU B8, fused (lazy_warp + fuse_model)", timed(fused), nbf, fl)


def cfg4():
    xa = [torch.randn(4, 128, 192, 384, device=dev), torch.randn(4, 256, 48, 96, device=dev)]
    xb = [torch.randn(4, 128, 192, 384, device=dev), torch.randn(4, 256, 48, 96, device=dev)]
    m = E.MatryoshkaDilatedCostVolumeList().to(dev)
    with torch.no_grad():
        ms = timed(lambda: m(xa, xb))
    report("cfg4 DCVNet Matryoshka list B4 384x768", ms, 240.0e6, 4.97e9)


def cfg5():
    """RAFT 1080p training path for ONE pair: build, 12 lookups, 12 lookup-backwards, pyramid backward."""
    f1 = torch.randn(1, 256, 135, 240, device=dev, requires_grad=True)
    f2 = torch.randn(1, 256, 135, 240, device=dev, requires_grad=True)
    grid = E.coords_grid(1, 135, 240, device=dev)
    coords = [(grid + torch.randn(1, 2, 135, 240, device=dev) * (2 + k)).clamp_(-16, 256) for k in range(12)]
    douts = [torch.randn(1, 324, 135, 240, device=dev) for _ in range(12)]

    def run():
        f1.grad = f2.grad = None
        fn = E.MutliScalePairwise4DCorr(f1, f2, num_levels=4, corr_radius=4)
        loss = 0
        for k in range(12):
            loss = loss + (fn(coords[k]) * douts[k]).sum()
        loss.backward()

    ms = timed(run, iters=3, warm=1)
    N = 32400
    nb = N * 42900 * 2 + 12 * 68.2e6 + 12 * 145.7e6 + 2 * N * 42900 * 4
    report("cfg5 RAFT 1080p fwd+bwd, 1 pair (build, 12 lookups, 12 lookup-bwd, pyramid-bwd)", ms, nb, 3 * 2.0 * N * N * 256)


def bwd():
    """Local-correlation adjoints (K4b) at the config 2 / config 3 shapes: forward + backward through autograd minus
    the forward alone (the loss glue is a single out.backward(dout))."""
    cases = [("cfg2 FlowNetC B8 256x48x64 P21 dp2", (8, 256, 48, 64), dict(patch_size=21, dilation_patch=2)),
             ("cfg3 finest PWC level B8 32x112x256 P9", (8, 32, 112, 256), dict(patch_size=9)),
             ("cfg3 PWC level B8 96x28x64 P9", (8, 96, 28, 64), dict(patch_size=9))]
    for name, shape, kw in cases:
        x1 = torch.randn(shape, device=dev, requires_grad=True)
        x2 = torch.randn(shape, device=dev, requires_grad=True)
        m = E.IterSpatialCorrelationSampler(kernel_size=1, padding=0, **kw)
        out = m(x1, x2)
        dout = torch.randn_like(out)
        with torch.no_grad():
            t_f = timed(lambda: m(x1.detach(), x2.detach()))

        def fb():
            x1.grad = x2.grad = None
            m(x1, x2).backward(dout)

        t_fb = timed(fb)
        nb = 2 * x1.numel() * 4 * 2 + out.numel() * 4
        report(f"K4b {name}: backward only (fwd {t_f:.3f} ms)", t_fb - t_f, nb, 2 * 2.0 * out.numel() * shape[1])


def upsample():
    """Convex flow upsampling at the headline size: 4 pairs, 135x240 -> 1080x1920 (called once per RAFT iteration);
    the reference's PyTorch ops (softmax + unfold + mul + sum + permute) beside it."""
    n, h, w, s = 4, 135, 240, 8
    flow = torch.randn(n, 2, h, w, device=dev) * 8
    logits = torch.randn(n, 9 * s * s, h, w, device=dev)
    nb = (logits.numel() + flow.numel() + n * 2 * s * s * h * w) * 4
    with torch.no_grad():
        report("convex upsample fwd B4 135x240 s8", timed(lambda: E.convex_upsample_flow(flow, logits, s)), nb, 0)

        def torch_ref():
            probs = torch.softmax(logits.view(n, 1, 9, s, s, h, w), dim=2)
            up = torch.nn.functional.unfold(flow, [3, 3], padding=1).view(n, 2, 9, 1, 1, h, w)
            return torch.sum(probs * up, dim=2).permute(0, 1, 4, 2, 5, 3).reshape(n, 2, s * h, s * w)

        report("  same, reference PyTorch ops on this GPU", timed(torch_ref), nb, 0)
    fl, lg = flow.clone().requires_grad_(True), logits.clone().requires_grad_(True)
    dout = torch.randn(n, 2, s * h, s * w, device=dev)

    def fb():
        fl.grad = lg.grad = None
        (E.convex_upsample_flow(fl, lg, s) * dout).sum().backward()

    report("convex upsample fwd+bwd B4 (incl. autograd glue)", timed(fb), 3 * nb, 0)


if __name__ == "__main__":
    which = sys.argv[1:] or ["cfg2", "cfg3", "cfg4", "cfg5", "upsample"]
    for w in which:
        globals()[w]()
