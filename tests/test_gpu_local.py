"""GPU parity: local displacement-window correlation (K4/K4b), CorrelationLayer, Matryoshka (K6), warp (K5).

Forward runs in both modes: 'tc' (tcgen05, fp16 operands / fp32 accumulation: <= 1e-3 relative, the north_star tolerance for
cost volumes) and 'exact' (fp32 CUDA cores: 2e-5, summation order only).  Backward kernels are fp32 in both modes."""

import numpy as np
import pytest
import torch

from gpu_util import assert_close, dev, to_gpu

pytestmark = pytest.mark.gpu

TOL = 2e-5
FWD_TOL = {"tc": 1e-3, "exact": 2e-5}
BWD_TOL = {"tc": 1e-3, "exact": 5e-5}  # adjoints: tensor-core path (fp16 operands) / exact fp32 kernels
SAMPLER = ["sampler_p9", "sampler_flownetc", "sampler_stride2", "sampler_aniso", "sampler_dcv_s4", "sampler_dil16"]


@pytest.fixture(params=["tc", "exact"])
def mode(request):
    import ezflow_b200 as E

    old = E.get_local_corr_mode()
    E.set_local_corr_mode(request.param)
    yield request.param
    E.set_local_corr_mode(old)


def _kw(g):
    kw = {}
    for k, v in g.items():
        if k.startswith("p_"):
            kw[k[2:]] = int(v) if v.ndim == 0 else tuple(int(t) for t in v)
    return kw


@pytest.mark.parametrize("name", SAMPLER)
def test_sampler_vs_reference_golden(golden, name, mode):
    import ezflow_b200 as E

    g = golden(name)
    kw = _kw(g)
    x1, x2 = to_gpu(g["x1"], True), to_gpu(g["x2"], True)
    mod = E.IterSpatialCorrelationSampler(**kw)
    out = mod(x1, x2)
    assert out.is_contiguous() and tuple(out.shape) == g["out"].shape
    assert_close(out, g["out"], FWD_TOL[mode], f"{name} fwd ({mode})")
    (out * to_gpu(g["dout"])).sum().backward()
    btol = BWD_TOL[mode]  # 'tc': the regular cases run their adjoints on tensor cores too (fp16 operands)
    assert_close(x1.grad, g["dx1"], btol, f"{name} dinput1 ({mode})")
    assert_close(x2.grad, g["dx2"], btol, f"{name} dinput2 ({mode})")
    # functional form, no grad
    with torch.no_grad():
        out2 = E.iter_spatial_correlation_sample(to_gpu(g["x1"]), to_gpu(g["x2"]), **kw)
    assert torch.equal(out2, out.detach())


def test_sampler_rejects_like_reference():
    import ezflow_b200 as E

    x = torch.zeros(1, 2, 4, 4, device=dev())
    for kw in (dict(kernel_size=3), dict(dilation=2), dict(patch_size=4)):
        with pytest.raises(NotImplementedError):
            E.IterSpatialCorrelationSampler(**kw)(x, x)
    with pytest.raises(RuntimeError):
        E.IterSpatialCorrelationSampler()(x.cpu(), x.cpu())


def test_view_like_callers():
    """FlowNetC / PWC call .view(B, -1, H, W) on the result (models/flownet_c.py:87-89, decoder/pyramid.py:101)."""
    import ezflow_b200 as E

    x = torch.randn(2, 16, 8, 12, device=dev())
    out = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=9, padding=0)(x, x)
    assert tuple(out.view(2, -1, 8, 12).shape) == (2, 81, 8, 12)


def test_corr_layer(golden, mode):
    import ezflow_b200 as E

    g = golden("corr_layer")
    out = E.CorrelationLayer()(to_gpu(g["x1"]), to_gpu(g["x2"]))
    assert_close(out, g["out"], FWD_TOL[mode], "CorrelationLayer default")
    out = E.CorrelationLayer(pad_size=3, max_displacement=3)(to_gpu(g["x1"]), to_gpu(g["x2"]))
    assert_close(out, g["out_p2_d3"], FWD_TOL[mode], "CorrelationLayer 3/3")


def test_matryoshka(golden, mode):
    import ezflow_b200 as E

    g = golden("matryoshka")
    x1, x2 = to_gpu(g["x1"]), to_gpu(g["x2"])
    m0 = E.MatryoshkaDilatedCostVolume().to(dev())
    assert m0.get_search_range() == 9
    np.testing.assert_array_equal(m0.get_relative_offsets().cpu().numpy(), g["offsets_default"])
    assert_close(m0(x1, x2), g["out_default"], FWD_TOL[mode], "matryoshka default")
    m1 = E.MatryoshkaDilatedCostVolume(num_groups=4, max_displacement=2, stride=2, dilations=[1, 3], use_relu=True).to(dev())
    assert_close(m1(x1, x2), g["out_g4"], FWD_TOL[mode], "matryoshka groups/relu")
    np.testing.assert_array_equal(m1.get_relative_offsets().cpu().numpy(), g["offsets_g4"])


def test_matryoshka_list(golden, mode):
    import ezflow_b200 as E

    g = golden("matryoshka_list")
    xa = [to_gpu(g["xa0"]), to_gpu(g["xa1"])]
    xb = [to_gpu(g["xb0"]), to_gpu(g["xb1"])]
    l0 = E.MatryoshkaDilatedCostVolumeList().to(dev())
    out = l0(xa, xb)
    assert out.is_contiguous()
    assert_close(out, g["out_default"], FWD_TOL[mode], "matryoshka list default")
    np.testing.assert_array_equal(l0.get_global_flow_offsets().cpu().numpy(), g["offsets_2d"])
    b, c, u, v, h, w = out.shape
    out.view(b, c * u * v, h, w)  # decoder/dilated_flow_stack_filter.py:132-133
    l1 = E.MatryoshkaDilatedCostVolumeList(normalize_feat_l2=True, use_relu=True).to(dev())
    assert_close(l1(xa, xb), g["out_norm_relu"], FWD_TOL[mode], "matryoshka list norm+relu")


def test_matryoshka_backward_vs_oracle(mode):
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    rng = np.random.default_rng(5)
    x1 = rng.standard_normal((1, 8, 9, 11)).astype(np.float32)
    x2 = rng.standard_normal((1, 8, 9, 11)).astype(np.float32)
    m = E.MatryoshkaDilatedCostVolume(num_groups=2, max_displacement=1, dilations=[1, 2]).to(dev())
    a, b = to_gpu(x1, True), to_gpu(x2, True)
    out = m(a, b)
    dout = rng.standard_normal(tuple(out.shape)).astype(np.float32)
    (out * to_gpu(dout)).sum().backward()
    d1 = np.zeros_like(x1).reshape(2, 4, 9, 11)
    d2 = np.zeros_like(d1)
    for i, d in enumerate([1, 2]):
        g = dout[:, i * 2 : (i + 1) * 2].reshape(2, 3, 3, 9, 11)
        e1, e2 = O.iter_spatial_correlation_sample_bwd(x1.reshape(2, 4, 9, 11), x2.reshape(2, 4, 9, 11), g, 3, 1, 0, d)
        d1 += e1
        d2 += e2
    assert_close(a.grad, d1.reshape(x1.shape), BWD_TOL[mode], f"matryoshka dx1 ({mode})")
    assert_close(b.grad, d2.reshape(x2.shape), BWD_TOL[mode], f"matryoshka dx2 ({mode})")


def test_warp(golden):
    import ezflow_b200 as E

    g = golden("warp")
    x, flow = to_gpu(g["x"], True), to_gpu(g["flow"], True)
    out = E.warp(x, flow)
    assert_close(out, g["out"], TOL, "warp fwd")
    (out * to_gpu(g["dout"])).sum().backward()
    assert_close(x.grad, g["dx"], TOL, "warp dx")
    assert_close(flow.grad, g["dflow"], 1e-4, "warp dflow")


@pytest.mark.parametrize("cfg", ["flownetc", "pwc", "dcv"])
def test_baseline_config_shapes_vs_oracle(cfg, mode):
    """BASELINE configs 2-4 at full spatial size (batch reduced so the numpy oracle stays in seconds)."""
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    rng = np.random.default_rng(17)
    if cfg == "flownetc":
        x1 = rng.standard_normal((1, 256, 48, 64)).astype(np.float32)
        x2 = rng.standard_normal((1, 256, 48, 64)).astype(np.float32)
        out = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=21, padding=0, dilation_patch=2)(to_gpu(x1), to_gpu(x2))
        ref = O.iter_spatial_correlation_sample(x1, x2, patch_size=21, dilation_patch=2)
        assert tuple(out.shape) == (1, 21, 21, 48, 64)
        assert_close(out, ref, FWD_TOL[mode], "cfg2")
    elif cfg == "pwc":
        for (c, h, w) in [(196, 7, 16), (128, 14, 32), (96, 28, 64), (64, 56, 128), (32, 112, 256)]:
            x1 = rng.standard_normal((2, c, h, w)).astype(np.float32)
            x2 = rng.standard_normal((2, c, h, w)).astype(np.float32)
            flow = (rng.standard_normal((2, 2, h, w)) * 2).astype(np.float32)
            wg = E.warp(to_gpu(x2), to_gpu(flow))
            wr = O.warp(x2, flow)
            assert_close(wg, wr, TOL, f"cfg3 warp {c}x{h}x{w}")
            out = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=9, padding=0)(to_gpu(x1), wg)
            ref = O.iter_spatial_correlation_sample(x1, wr, patch_size=9)
            assert_close(out, ref, max(5e-5, FWD_TOL[mode]), f"cfg3 corr {c}x{h}x{w}")
    else:
        xa = [rng.standard_normal((1, 128, 192, 384)).astype(np.float32), rng.standard_normal((1, 256, 48, 96)).astype(np.float32)]
        xb = [rng.standard_normal((1, 128, 192, 384)).astype(np.float32), rng.standard_normal((1, 256, 48, 96)).astype(np.float32)]
        m = E.MatryoshkaDilatedCostVolumeList().to(dev())
        out = m([to_gpu(t) for t in xa], [to_gpu(t) for t in xb])
        assert tuple(out.shape) == (1, 7, 9, 9, 48, 96)
        ref = O.matryoshka_cost_volume_list(xa, xb)
        assert_close(out, ref, FWD_TOL[mode], "cfg4")


@pytest.mark.parametrize("cfg", [(1, 4, 3, 3, 3, 1, 0, 1), (1, 7, 5, 9, 5, 2, 1, 2), (2, 130, 6, 6, 9, 1, 4, 1), (1, 16, 20, 30, 3, 9, 0, 1),
                                 (1, 16, 20, 30, 41, 1, 0, 1), (1, 24, 18, 26, 33, 1, 0, 2)])
def test_tiny_and_ragged_shapes_vs_oracle(cfg, mode):
    """Maps smaller than one query tile, odd channel counts (padded to 64), padding with stride, stride > 8; windows wider than
    32 columns (the out-of-image zeros are written from a 32-bit column mask per pass: two passes), most of them outside the map."""
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    B, C, H, W, P, s, pad, dp = cfg
    rng = np.random.default_rng(200 + C)
    x1 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    x2 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    ref = O.iter_spatial_correlation_sample(x1, x2, patch_size=P, stride=s, padding=pad, dilation_patch=dp)
    out = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=P, stride=s, padding=pad, dilation_patch=dp)(to_gpu(x1), to_gpu(x2))
    assert tuple(out.shape) == ref.shape
    assert_close(out, ref, FWD_TOL[mode], f"{cfg} ({mode})")


def test_empty_batch_returns_empty_outputs():
    """B = 0 (e.g. the tail of a DataParallel scatter): every op returns an empty tensor of the reference's shape."""
    import ezflow_b200 as E

    z = torch.zeros(0, 16, 12, 20, device=dev())
    fl = torch.zeros(0, 2, 12, 20, device=dev())
    assert tuple(E.IterSpatialCorrelationSampler(patch_size=9)(z, z).shape) == (0, 9, 9, 12, 20)
    assert tuple(E.CorrelationLayer()(z, z).shape) == (0, 81, 12, 20)
    assert tuple(E.MatryoshkaDilatedCostVolume()(z, z).shape) == (0, 6, 9, 9, 12, 20)
    assert tuple(E.warp(z, fl).shape) == (0, 16, 12, 20)
    assert tuple(E.convex_upsample_flow(fl, torch.zeros(0, 576, 12, 20, device=dev()), 8).shape) == (0, 2, 96, 160)
    fn = E.MutliScalePairwise4DCorr(z, z, num_levels=2, corr_radius=3)
    assert tuple(fn(fl).shape) == (0, 2 * 49, 12, 20)
    assert [tuple(t.shape) for t in fn.corr_pyramid] == [(0, 1, 12, 20), (0, 1, 6, 10)]  # reference pairwise.py:31-41 on B = 0


def test_matryoshka_relu_backward_at_exact_zero_costs():
    """use_relu=True where in-image costs are EXACTLY zero (all-zero feature pixels): F.leaky_relu's backward applies the
    slope at 0 (x > 0 selects the identity branch), reference learnable_cost.py:322-323."""
    import ezflow_b200 as E
    import torch.nn.functional as F

    rng = np.random.default_rng(9)
    x1 = rng.standard_normal((1, 8, 9, 11)).astype(np.float32)
    x2 = rng.standard_normal((1, 8, 9, 11)).astype(np.float32)
    x1[:, :, 2:5, 3:7] = 0.0  # every cost of these queries is exactly 0
    x2[:, :, 6, :] = 0.0
    m = E.MatryoshkaDilatedCostVolume(num_groups=2, max_displacement=1, dilations=[1, 2], use_relu=True).to(dev())
    old = E.get_local_corr_mode()
    E.set_local_corr_mode("exact")
    try:
        a, b = to_gpu(x1, True), to_gpu(x2, True)
        out = m(a, b)
        dout = to_gpu(rng.standard_normal(tuple(out.shape)).astype(np.float32))
        (out * dout).sum().backward()
        # comparator: the un-activated volume from the same kernels, then PyTorch's own leaky_relu (forward and backward)
        m_lin = E.MatryoshkaDilatedCostVolume(num_groups=2, max_displacement=1, dilations=[1, 2], use_relu=False).to(dev())
        a2, b2 = to_gpu(x1, True), to_gpu(x2, True)
        ref = F.leaky_relu(m_lin(a2, b2), 0.1)
        assert (ref == 0).sum().item() > 100
        (ref * dout).sum().backward()
        assert_close(out, ref, 2e-5, "matryoshka relu fwd")
        assert_close(a.grad, a2.grad, 2e-5, "matryoshka relu dx1 at zero costs")
        assert_close(b.grad, b2.grad, 2e-5, "matryoshka relu dx2 at zero costs")
    finally:
        E.set_local_corr_mode(old)


def test_correlation_layer_pad_larger_than_displacement(mode):
    """pad_size > max_displacement: the reference (layer.py:44-60) slices features2 padded by pad_size with offsets 0..2d,
    an off-centre window; restated here with the same slices in PyTorch."""
    import ezflow_b200 as E
    import torch.nn.functional as F

    rng = np.random.default_rng(4)
    x1 = to_gpu(rng.standard_normal((2, 24, 10, 13)).astype(np.float32))
    x2 = to_gpu(rng.standard_normal((2, 24, 10, 13)).astype(np.float32))
    pad, d = 5, 2
    H, W = x1.shape[-2:]
    x2p = F.pad(x2, (pad, pad, pad, pad))
    ref = torch.cat([(x1 * x2p[:, :, dy : dy + H, dx : dx + W]).mean(1, keepdim=True) for dy in range(2 * d + 1) for dx in range(2 * d + 1)], 1)
    out = E.CorrelationLayer(pad_size=pad, max_displacement=d)(x1, x2)
    assert_close(out, ref, FWD_TOL[mode], "CorrelationLayer pad 5 / displacement 2")
    with pytest.raises(NotImplementedError):
        E.CorrelationLayer(pad_size=1, max_displacement=2)(x1, x2)


def _torch_sampler(x1, x2, patch, stride=1, padding=0, dp=1):
    """fp32 PyTorch restatement of iter_spatial_correlation_sample (reference sampler.py:96-129: pad, one shifted slice per
    displacement, multiply, sum over channels) — the full-size checker where the numpy oracle would take minutes."""
    import torch.nn.functional as F

    b, c, h, w = x1.shape
    a = F.pad(x1, (padding, padding, padding, padding))
    t = F.pad(x2, (padding, padding, padding, padding))
    md = dp * (patch - 1) // 2
    hp, wp = a.shape[2:]
    t = F.pad(t, (md, md, md, md))
    a = a[:, :, ::stride, ::stride]
    out = torch.empty(b, patch, patch, a.shape[2], a.shape[3], device=x1.device)
    for i in range(patch):
        for j in range(patch):
            out[:, i, j] = (a * t[:, :, i * dp : i * dp + hp : stride, j * dp : j * dp + wp : stride]).sum(1)
    return out


@pytest.mark.parametrize("cfg", ["cfg2", "cfg3_fine", "cfg4_s8_dil", "cfg4_s2"])
def test_baseline_configs_full_batch(cfg, mode):
    """BASELINE configs 2-4 at their FULL batch sizes (many work items per persistent CTA, nsplit, lattice mode, stride 4)
    against the PyTorch fp32 restatement on the same GPU."""
    import ezflow_b200 as E

    g = torch.Generator(device=dev()).manual_seed(23)
    shape, kw = {
        "cfg2": ((8, 256, 48, 64), dict(patch=21, dp=2)),
        "cfg3_fine": ((8, 32, 112, 256), dict(patch=9)),
        "cfg4_s8_dil": ((4, 256, 48, 96), dict(patch=9, dp=5)),
        "cfg4_s2": ((4, 128, 192, 384), dict(patch=9, stride=4)),
    }[cfg]
    x1 = torch.randn(shape, device=dev(), generator=g)
    x2 = torch.randn(shape, device=dev(), generator=g)
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        ref = _torch_sampler(x1, x2, kw["patch"], kw.get("stride", 1), 0, kw.get("dp", 1))
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    m = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=kw["patch"], stride=kw.get("stride", 1), padding=0,
                                        dilation_patch=kw.get("dp", 1))
    out = m(x1, x2)
    assert tuple(out.shape) == tuple(ref.shape)
    assert_close(out, ref, FWD_TOL[mode] if mode == "tc" else 5e-5, f"{cfg} full batch")
    # run-to-run determinism at full size (persistent CTAs must not race on the shared rows / accumulators)
    assert torch.equal(out, m(x1, x2))


@pytest.mark.parametrize("cfg", ["cfg2", "cfg3_fine", "pad_dil"])
def test_adjoints_full_size_vs_torch_autograd(cfg, mode):
    """K4b at BASELINE shapes (tiled stride-1 kernels, many tiles / channel chunks) against autograd of the PyTorch restatement."""
    import ezflow_b200 as E

    g = torch.Generator(device=dev()).manual_seed(29)
    shape, kw = {
        "cfg2": ((4, 256, 48, 64), dict(patch=21, dp=2)),
        "cfg3_fine": ((8, 32, 112, 256), dict(patch=9)),
        "pad_dil": ((2, 40, 37, 45), dict(patch=5, dp=3, padding=4)),  # ragged tiles, 40 channels (2.5 chunks), padding
    }[cfg]
    x1 = torch.randn(shape, device=dev(), generator=g)
    x2 = torch.randn(shape, device=dev(), generator=g)
    a1, a2 = x1.clone().requires_grad_(True), x2.clone().requires_grad_(True)
    ref = _torch_sampler(a1, a2, kw["patch"], 1, kw.get("padding", 0), kw.get("dp", 1))
    dout = torch.randn(ref.shape, device=dev(), generator=g)
    ref.backward(dout)
    b1, b2 = x1.clone().requires_grad_(True), x2.clone().requires_grad_(True)
    m = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=kw["patch"], padding=kw.get("padding", 0), dilation_patch=kw.get("dp", 1))
    m(b1, b2).backward(dout)
    btol = BWD_TOL[mode] if cfg == "cfg2" else 5e-5  # cfg2: tensor-core adjoint in 'tc' mode; the other two take the fp32 kernels
    assert_close(b1.grad, a1.grad, btol, f"{cfg} d input1 ({mode})")
    assert_close(b2.grad, a2.grad, btol, f"{cfg} d input2 ({mode})")


@pytest.mark.parametrize("cfg", [
    dict(shape=(2, 8, 13, 18), patch=9, dp=1),      # ragged map, two query tiles per row, image borders everywhere
    dict(shape=(1, 16, 12, 20), patch=21, dp=1),    # window wider than the map (4x40 target tiles)
    dict(shape=(1, 40, 19, 37), patch=5, dp=2),     # lattice mode with ragged residue classes, C not a multiple of 64
    dict(shape=(2, 256, 9, 17), patch=3, dp=1),     # full 256 channels (4 chunks)
    dict(shape=(1, 24, 48, 64), patch=7, dp=3),     # lattice 3: classes of different sizes
])
def test_tensor_core_adjoint_vs_oracle(cfg):
    """K4b on tcgen05 called directly through the C ABI (ezb_local_corr_bwd_ex with a workspace), whatever the size heuristic of
    the host class would choose: both adjoints against the numpy oracle, 1e-3 (fp16 operands); then the same call without a
    workspace (fp32 kernels) at 5e-5."""
    import ctypes

    from ezflow_b200 import _cabi
    from oracle import similarity_oracle as O

    rng = np.random.default_rng(sum(cfg["shape"]) + cfg["patch"])
    B, C, H, W = cfg["shape"]
    P, dp = cfg["patch"], cfg["dp"]
    x1 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    x2 = (rng.standard_normal((B, C, H, W)) * 3.0).astype(np.float32)
    dout = (rng.standard_normal((B, P, P, H, W)) * 0.01).astype(np.float32)
    d1, d2 = O.iter_spatial_correlation_sample_bwd(x1, x2, dout, patch_size=P, stride=1, padding=0, dilation_patch=dp)
    lib = _cabi.lib()
    a, b, g = to_gpu(x1), to_gpu(x2), to_gpu(dout)
    n = ctypes.c_size_t(0)
    _cabi.check(lib.ezb_local_corr_bwd_workspace_bytes(B, C, H, W, ctypes.byref(n)))
    for ws_bytes, tol in ((n.value, 1e-3), (0, 5e-5)):
        ws = torch.empty(max(ws_bytes, 1), dtype=torch.uint8, device=dev())
        o1, o2 = torch.full_like(a, float("nan")), torch.full_like(b, float("nan"))  # every element must be written
        _cabi.check(lib.ezb_local_corr_bwd_ex(_cabi.ptr(a), _cabi.ptr(b), _cabi.ptr(g), B, C, H, W, P, P, 1, 1, 0, 0, dp, dp, 1.0, 0,
                                              _cabi.ptr(o1), _cabi.ptr(o2), _cabi.ptr(ws) if ws_bytes else None, ws_bytes,
                                              _cabi.stream_ptr(dev())))
        assert_close(o1, d1, tol, f"{cfg} din1 ({'tensor cores' if ws_bytes else 'fp32'})")
        assert_close(o2, d2, tol, f"{cfg} din2 ({'tensor cores' if ws_bytes else 'fp32'})")
    # one gradient only
    o1 = torch.full_like(a, float("nan"))
    ws = torch.empty(n.value, dtype=torch.uint8, device=dev())
    _cabi.check(lib.ezb_local_corr_bwd_ex(_cabi.ptr(a), _cabi.ptr(b), _cabi.ptr(g), B, C, H, W, P, P, 1, 1, 0, 0, dp, dp, 0.5, 0,
                                          _cabi.ptr(o1), None, _cabi.ptr(ws), n.value, _cabi.stream_ptr(dev())))
    assert_close(o1, 0.5 * d1, 1e-3, f"{cfg} din1 only, scale 0.5")


@pytest.mark.parametrize("cfg", [
    dict(shape=(2, 16, 26, 50), md=2, dil=[1, 2, 3], relu=False),   # every dilation on the tensor-core adjoint ('tc' mode)
    dict(shape=(1, 24, 32, 48), md=4, dil=[1, 9], relu=True),       # dilation 9 leaves too few pixels per residue class: fp32 kernels + add
    dict(shape=(1, 8, 9, 11), md=1, dil=[2, 1], relu=False),        # tiny map: dilation 2 on the fp32 kernels first, dilation 1 accumulates
])
def test_matryoshka_backward_single_group_all_dilations_in_one_call(cfg, mode):
    """ezb_matryoshka_bwd (one C call for all dilations, G = 1 as in DCVNet) against the oracle's per-dilation adjoints."""
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    rng = np.random.default_rng(sum(cfg["shape"]))
    x1 = rng.standard_normal(cfg["shape"]).astype(np.float32)
    x2 = rng.standard_normal(cfg["shape"]).astype(np.float32)
    P = 2 * cfg["md"] + 1
    m = E.MatryoshkaDilatedCostVolume(num_groups=1, max_displacement=cfg["md"], dilations=cfg["dil"], use_relu=cfg["relu"]).to(dev())
    a, b = to_gpu(x1, True), to_gpu(x2, True)
    out = m(a, b)
    dout = rng.standard_normal(tuple(out.shape)).astype(np.float32)
    (out * to_gpu(dout)).sum().backward()
    # the activation mask is taken from the forward that actually ran: at a cost within fp16 rounding of zero the tensor-core
    # forward and the oracle may sit on different sides of the LeakyReLU kink, which is not what this test is about
    mask = out.detach().cpu().numpy() > 0
    g = dout * np.where(mask, 1.0, 0.1).astype(np.float32) if cfg["relu"] else dout
    d1, d2 = np.zeros_like(x1), np.zeros_like(x2)
    for i, d in enumerate(cfg["dil"]):
        e1, e2 = O.iter_spatial_correlation_sample_bwd(x1, x2, np.ascontiguousarray(g[:, i]), P, 1, 0, d)
        d1 += e1
        d2 += e2
    assert_close(a.grad, d1, BWD_TOL[mode], f"{cfg} dx1 ({mode})")
    assert_close(b.grad, d2, BWD_TOL[mode], f"{cfg} dx2 ({mode})")


def test_few_surviving_dot_products_bound_against_their_scale():
    """Regression for a fuzz case (B=2, C=130, 4x5 maps, P=5, stride 4, padding 1, dilation_patch 3: 200 outputs of which most
    point outside the image): the tensor-core path's error measured against max|ref| was 1.19e-3 — with a handful of
    random-sign dot products that statistic depends on how small the largest survivor happens to be.  The bound that holds
    for fp16 operands is against the scale of a C-term dot product: |err| <= 1e-3 * sqrt(C) * rms(x1) * rms(x2)."""
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    B, C, H, W, P, s, pad, dp = 2, 130, 4, 5, 5, 4, 1, 3
    for seed in range(6):
        rng = np.random.default_rng(seed)
        x1 = rng.standard_normal((B, C, H, W)).astype(np.float32)
        x2 = (rng.standard_normal((B, C, H, W)) * 10.0 ** rng.uniform(-2, 2)).astype(np.float32)
        ref = O.iter_spatial_correlation_sample(x1, x2, patch_size=P, stride=s, padding=pad, dilation_patch=dp)
        old = E.get_local_corr_mode()
        E.set_local_corr_mode("tc")
        try:
            out = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=P, stride=s, padding=pad, dilation_patch=dp)(to_gpu(x1), to_gpu(x2))
        finally:
            E.set_local_corr_mode(old)
        out = out.cpu().numpy()
        scale = float(np.sqrt(C) * np.sqrt((x1.astype(np.float64) ** 2).mean()) * np.sqrt((x2.astype(np.float64) ** 2).mean()))
        err = float(np.abs(out.astype(np.float64) - ref).max())
        assert err <= 1e-3 * scale, f"seed {seed}: max abs error {err:.3e} vs dot-product scale {scale:.3e}"
        assert np.array_equal(out == 0, ref == 0), "displacements outside the image must be exact zeros"


def _torch_matryoshka(x1, x2, groups, md, stride, dilations, relu):
    """fp32 PyTorch restatement of MatryoshkaDilatedCostVolume.forward (reference learnable_cost.py:306-325) on the GPU."""
    b, c, h, w = x1.shape
    a = x1.reshape(b * groups, c // groups, h, w)
    t = x2.reshape(b * groups, c // groups, h, w)
    outs = []
    for d in dilations:
        o = _torch_sampler(a, t, 2 * md + 1, stride, 0, d)
        outs.append(o.reshape(b, groups, 2 * md + 1, 2 * md + 1, o.shape[-2], o.shape[-1]))
    out = torch.cat(outs, dim=1)
    return torch.nn.functional.leaky_relu(out, 0.1) if relu else out


@pytest.mark.parametrize(
    "cfg",
    [
        ((4, 256, 48, 96), 1, 4, 1, [1, 2, 3, 5, 9, 16], False),   # config 4's stride-1 scale at its full batch
        ((2, 96, 37, 53), 2, 4, 1, [1, 2, 3, 5, 9, 16], True),     # ragged map (partial query / target tiles), groups, LeakyReLU
        ((1, 64, 40, 72), 1, 2, 1, [1, 4, 7, 40], False),          # a dilation whose margin exceeds the sparse-sweep tables
        ((2, 64, 33, 47), 1, 3, 2, [1, 2, 4, 6, 11], True),        # stride 2 with five dilations
        ((1, 32, 24, 40), 4, 1, 1, [1, 2, 3, 4, 5, 6, 7, 8], False),  # eight dilations (every table bit), 3x3 window
    ],
)
def test_many_dilations_sparse_sweep_vs_torch(cfg, mode):
    """K6 with D >= 4: the sparse sweep (tiles no displacement reaches are skipped by all three roles), (tile, dilation) pairs
    dealt to the epilogue warps, zeros written by the idle warps — against the PyTorch fp32 restatement of the reference's loop
    (learnable_cost.py:306-325), full batch, ragged borders, table overflow; and run-to-run determinism."""
    import ezflow_b200 as E

    shape, groups, md, stride, dilations, relu = cfg
    g = torch.Generator(device=dev()).manual_seed(41)
    x1 = torch.randn(shape, device=dev(), generator=g)
    x2 = torch.randn(shape, device=dev(), generator=g)
    old = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
    try:
        ref = _torch_matryoshka(x1, x2, groups, md, stride, dilations, relu)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    m = E.MatryoshkaDilatedCostVolume(num_groups=groups, max_displacement=md, stride=stride, dilations=dilations, use_relu=relu).to(dev())
    out = m(x1, x2)
    assert tuple(out.shape) == tuple(ref.shape)
    assert_close(out, ref, FWD_TOL[mode] if mode == "tc" else 5e-5, f"{cfg} ({mode})")
    assert torch.equal(out, m(x1, x2))
