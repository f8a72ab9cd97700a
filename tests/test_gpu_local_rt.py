"""GPU parity: the register-tiled fp32 kernel for PWC-Net-shaped local correlations (csrc/local_corr_rt.cu: stride 1, no
padding, no dilation, odd patch <= 9), forced on and off through EZB_LOCAL_CORR_RT so that both it and the tensor-core path
stay covered whatever the library's own choice is.  fp32 throughout: 2e-5 against the numpy oracle (summation order only)."""

import os

import numpy as np
import pytest
import torch

from gpu_util import assert_close, dev, to_gpu

pytestmark = pytest.mark.gpu


def _setenv(**kv):
    old = {k: os.environ.get(k) for k in kv}
    for k, v in kv.items():
        if v is None:
            os.environ.pop(k, None)
        else:
            os.environ[k] = v
    return old


@pytest.fixture(params=["large_map_tiling", "small_map_tiling"])
def force_rt(request):
    """Force the register-tiled kernel, in its <8 rows, 3 displacement rows per thread> and its <4, 1> configuration."""
    old = _setenv(EZB_LOCAL_CORR_RT="1", EZB_LOCAL_CORR_RT_SMALL="1" if request.param == "small_map_tiling" else "0")
    yield request.param
    _setenv(**old)


@pytest.fixture(params=["tc", "exact"])
def mode(request):
    import ezflow_b200 as E

    old = E.get_local_corr_mode()
    E.set_local_corr_mode(request.param)
    yield request.param
    E.set_local_corr_mode(old)


# (B, C, H, W, P): vector path (W % 4 == 0) and scalar path, ragged tiles, channel counts off the 8-channel stage, every patch size
SHAPES = [
    (2, 32, 16, 64, 9), (1, 13, 9, 36, 9), (2, 5, 11, 30, 9), (1, 64, 8, 32, 9), (1, 3, 1, 4, 9), (2, 17, 19, 45, 7),
    (1, 8, 10, 40, 7), (2, 9, 7, 33, 5), (1, 24, 17, 68, 5), (1, 6, 12, 20, 3), (2, 10, 5, 7, 3), (1, 130, 9, 16, 9),
]


@pytest.mark.parametrize("shape", SHAPES)
def test_register_tiled_forward_vs_oracle(shape, mode, force_rt):
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    B, C, H, W, P = shape
    rng = np.random.default_rng(hash(shape) % (1 << 31))
    x1 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    x2 = (rng.standard_normal((B, C, H, W)) * 3).astype(np.float32)
    ref = O.iter_spatial_correlation_sample(x1, x2, patch_size=P)
    out = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=P)(to_gpu(x1), to_gpu(x2))
    assert tuple(out.shape) == ref.shape and out.is_contiguous()
    assert_close(out, ref, 2e-5, f"register-tiled {shape} ({mode})")  # exact fp32 in both modes


@pytest.mark.parametrize("shape", [(2, 32, 16, 64, 9), (1, 13, 9, 37, 9), (2, 9, 7, 33, 5)])
def test_register_tiled_warp_fused_vs_oracle(shape, force_rt):
    """warp -> correlation -> LeakyReLU in one pass (reference decoder/pyramid.py:98-102,138-141): the halo is warped while staged."""
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    B, C, H, W, P = shape
    rng = np.random.default_rng(11)
    x1 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    x2 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    flow = (rng.standard_normal((B, 2, H, W)) * 3).astype(np.float32)
    ref = O.leaky_relu(O.iter_spatial_correlation_sample(x1, O.warp(x2, flow), patch_size=P), 0.1)
    m = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=P)
    m.fused_negative_slope = 0.1
    out = m(to_gpu(x1), E.lazy_warp(to_gpu(x2), to_gpu(flow)))
    assert_close(out, ref, 5e-5, f"register-tiled warp-fused {shape}")


def test_register_tiled_matches_tensor_core_path_on_pwc_level():
    """The library's own choice (register-tiled for the two finest PWC-Net levels) against the tensor-core path forced on, and
    against fp32 PyTorch ops, at the full (8, 32, 112, 256) level."""
    import ezflow_b200 as E
    from test_gpu_local import _torch_sampler

    g = torch.Generator(device=dev()).manual_seed(5)
    x1 = torch.randn((8, 32, 112, 256), device=dev(), generator=g)
    x2 = torch.randn((8, 32, 112, 256), device=dev(), generator=g)
    m = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=9)
    ref = _torch_sampler(x1, x2, 9)
    old = _setenv(EZB_LOCAL_CORR_RT=None, EZB_LOCAL_CORR_RT_SMALL=None)
    try:
        auto = m(x1, x2)
        os.environ["EZB_LOCAL_CORR_RT"] = "0"
        tc = m(x1, x2)
        os.environ["EZB_LOCAL_CORR_RT"] = "1"
        rt = m(x1, x2)
    finally:
        _setenv(**old)
    assert_close(rt, ref, 5e-5, "register-tiled vs fp32 torch ops")
    assert_close(tc, ref, 1e-3, "tensor-core path vs fp32 torch ops")
    assert torch.equal(auto, rt), "the library should pick the register-tiled kernel for this level"


def test_register_tiled_gradients_flow(force_rt):
    """Backward is unchanged (it never depended on which forward kernel ran); checked once through autograd."""
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    rng = np.random.default_rng(3)
    x1 = rng.standard_normal((1, 12, 9, 20)).astype(np.float32)
    x2 = rng.standard_normal((1, 12, 9, 20)).astype(np.float32)
    dout = rng.standard_normal((1, 9, 9, 9, 20)).astype(np.float32)
    t1, t2 = to_gpu(x1, True), to_gpu(x2, True)
    old = E.get_local_corr_mode()
    E.set_local_corr_mode("exact")
    try:
        E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=9)(t1, t2).backward(to_gpu(dout))
    finally:
        E.set_local_corr_mode(old)
    d1, d2 = O.iter_spatial_correlation_sample_bwd(x1, x2, dout, patch_size=9)
    assert_close(t1.grad, d1, 5e-5, "dinput1")
    assert_close(t2.grad, d2, 5e-5, "dinput2")


# (B, C, H, W, P): W % 4 == 0 (the register-tiled adjoints' vector path); ragged tiles, channel tails, every patch size
BWD_SHAPES = [(2, 32, 16, 64, 9), (1, 13, 9, 36, 9), (1, 40, 11, 44, 9), (2, 17, 19, 48, 7), (2, 9, 7, 32, 5), (1, 6, 12, 20, 3), (1, 3, 1, 4, 9)]


@pytest.mark.parametrize("shape", BWD_SHAPES)
def test_register_tiled_adjoints_vs_oracle(shape):
    """din1 / din2 through local_corr_bwd_rt_kernel (the 'exact' path for stride 1, no dilation, odd window <= 9) vs the oracle,
    and bit-for-bit reproducible; the round-1 tiled kernels (EZB_LOCAL_CORR_BWD_RT=0) must agree to summation order."""
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    B, C, H, W, P = shape
    rng = np.random.default_rng(P * 1000 + C)
    x1 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    x2 = (rng.standard_normal((B, C, H, W)) * 2).astype(np.float32)
    dout = rng.standard_normal((B, P, P, H, W)).astype(np.float32)
    d1, d2 = O.iter_spatial_correlation_sample_bwd(x1, x2, dout, patch_size=P)
    old_mode = E.get_local_corr_mode()
    E.set_local_corr_mode("exact")
    grads = {}
    try:
        for flag in ("1", "0"):
            old = _setenv(EZB_LOCAL_CORR_BWD_RT=flag)
            try:
                t1, t2 = to_gpu(x1, True), to_gpu(x2, True)
                E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=P)(t1, t2).backward(to_gpu(dout))
                grads[flag] = (t1.grad.clone(), t2.grad.clone())
            finally:
                _setenv(**old)
    finally:
        E.set_local_corr_mode(old_mode)
    for flag, (g1, g2) in grads.items():
        assert_close(g1, d1, 5e-5, f"dinput1 {shape} rt={flag}")
        assert_close(g2, d2, 5e-5, f"dinput2 {shape} rt={flag}")


def test_register_tiled_adjoints_only_one_gradient():
    """Only input2 requires grad (the other launch is skipped)."""
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    rng = np.random.default_rng(8)
    x1 = rng.standard_normal((1, 20, 10, 24)).astype(np.float32)
    x2 = rng.standard_normal((1, 20, 10, 24)).astype(np.float32)
    dout = rng.standard_normal((1, 9, 9, 10, 24)).astype(np.float32)
    old_mode = E.get_local_corr_mode()
    E.set_local_corr_mode("exact")
    try:
        t2 = to_gpu(x2, True)
        E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=9)(to_gpu(x1), t2).backward(to_gpu(dout))
    finally:
        E.set_local_corr_mode(old_mode)
    assert_close(t2.grad, O.iter_spatial_correlation_sample_bwd(x1, x2, dout, patch_size=9)[1], 5e-5, "dinput2 only")


# (B, C, H, W, P, S): stride 2 / 4, W % 4 == 0 and ceil(W / S) % 4 == 0 (the strided kernel's vector path); ragged tiles in both
# directions, channel counts off the 4-channel stage, maps whose last query row / column is not a multiple of the stride
STRIDED_SHAPES = [
    (2, 16, 32, 64, 9, 4), (1, 13, 19, 48, 9, 4), (1, 5, 70, 144, 9, 4), (2, 7, 9, 16, 9, 4), (1, 130, 16, 32, 9, 4),
    (2, 12, 18, 40, 9, 2), (1, 9, 33, 72, 9, 2), (1, 6, 21, 32, 7, 4), (1, 10, 14, 24, 5, 2), (2, 4, 12, 16, 3, 4),
    (1, 8, 17, 62 * 4, 9, 4),
]


@pytest.mark.parametrize("shape", STRIDED_SHAPES)
def test_strided_register_tiled_forward_vs_oracle(shape, mode):
    """local_corr_rt_strided_kernel (DCVNet's fine scale: stride-2 features sampled every 4 pixels, learnable_cost.py:420-434)
    through the sampler class, in both modes (the kernel is exact fp32), and switched off (EZB_LOCAL_CORR_RT=0: the generic
    path) on the same inputs."""
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    B, C, H, W, P, S = shape
    rng = np.random.default_rng(hash(shape) % (1 << 31))
    x1 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    x2 = (rng.standard_normal((B, C, H, W)) * 2).astype(np.float32)
    ref = O.iter_spatial_correlation_sample(x1, x2, patch_size=P, stride=S)
    m = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=P, stride=S)
    out = m(to_gpu(x1), to_gpu(x2))
    assert tuple(out.shape) == ref.shape and out.is_contiguous()
    covered = W % 4 == 0 and -(-W // S) % 4 == 0  # otherwise the generic path runs (fp16 operands in 'tc' mode)
    assert_close(out, ref, 2e-5 if covered or mode == "exact" else 1e-3, f"strided register-tiled {shape} ({mode})")
    old = _setenv(EZB_LOCAL_CORR_RT="0")
    try:
        other = m(to_gpu(x1), to_gpu(x2))
    finally:
        _setenv(**old)
    assert_close(other, ref, 1e-3 if mode == "tc" else 2e-5, f"generic path {shape} ({mode})")


@pytest.mark.parametrize("use_relu", [False, True])
def test_strided_register_tiled_in_matryoshka_list(use_relu):
    """The DCVNet cost-volume list (fine scale at stride 4 -> strided kernel, coarse scale -> band GEMM) into one concatenated
    volume, groups > 1 and the fused LeakyReLU included."""
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    rng = np.random.default_rng(21)
    x1 = [rng.standard_normal((2, 8, 32, 64)).astype(np.float32), rng.standard_normal((2, 16, 8, 16)).astype(np.float32)]
    x2 = [rng.standard_normal((2, 8, 32, 64)).astype(np.float32), rng.standard_normal((2, 16, 8, 16)).astype(np.float32)]
    ref = O.matryoshka_cost_volume_list(x1, x2, num_groups=2, max_displacement=4, encoder_output_strides=(2, 8),
                                        dilations=((1,), (1, 2, 3)), normalize_feat_l2=False, use_relu=use_relu)
    m = E.MatryoshkaDilatedCostVolumeList(num_groups=2, max_displacement=4, encoder_output_strides=[2, 8], dilations=[[1], [1, 2, 3]],
                                          normalize_feat_l2=False, use_relu=use_relu).to(dev())
    out = m([to_gpu(a) for a in x1], [to_gpu(a) for a in x2])
    assert tuple(out.shape) == ref.shape
    assert_close(out, ref, 1e-3, f"matryoshka list use_relu={use_relu}")
    fine = out[:, :2]  # the stride-4 scale's groups: exact fp32
    assert_close(fine, ref[:, :2], 2e-5, "fine scale (strided register-tiled kernel)")


# (B, C, H, W, P): stride 4, W % 4 == 0; H not a multiple of 4, single-row / single-column query grids, channel counts that
# do not divide into the launch's channel chunks, every patch size
STRIDED_BWD_SHAPES = [(2, 16, 32, 64, 9), (1, 13, 19, 48, 9), (1, 5, 70, 144, 9), (2, 7, 9, 16, 9), (1, 3, 4, 4, 9), (1, 6, 21, 32, 7),
                      (2, 10, 14, 24, 5), (1, 4, 12, 16, 3), (1, 37, 6, 132, 9)]


@pytest.mark.parametrize("shape", STRIDED_BWD_SHAPES)
def test_strided_adjoints_vs_oracle(shape):
    """local_corr_bwd_strided_kernel (DCVNet's fine scale, stride 4) vs the oracle's closed-form adjoint; the round-1 gather
    kernels (EZB_LOCAL_CORR_BWD_RT=0) must agree to summation order; din1 is exactly zero off the query lattice."""
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    B, C, H, W, P = shape
    rng = np.random.default_rng(P * 100 + C)
    x1 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    x2 = (rng.standard_normal((B, C, H, W)) * 2).astype(np.float32)
    oh, ow = -(-H // 4), -(-W // 4)
    dout = rng.standard_normal((B, P, P, oh, ow)).astype(np.float32)
    d1, d2 = O.iter_spatial_correlation_sample_bwd(x1, x2, dout, patch_size=P, stride=4)
    grads = {}
    for flag in ("1", "0"):
        old = _setenv(EZB_LOCAL_CORR_BWD_RT=flag)
        try:
            t1, t2 = to_gpu(x1, True), to_gpu(x2, True)
            E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=P, stride=4)(t1, t2).backward(to_gpu(dout))
            grads[flag] = (t1.grad.clone(), t2.grad.clone())
        finally:
            _setenv(**old)
    for flag, (g1, g2) in grads.items():
        assert_close(g1, d1, 5e-5, f"dinput1 {shape} rt={flag}")
        assert_close(g2, d2, 5e-5, f"dinput2 {shape} rt={flag}")
    g1 = grads["1"][0]
    lattice = torch.zeros_like(g1, dtype=torch.bool)
    lattice[:, :, ::4, ::4] = True
    assert float(g1[~lattice].abs().max()) == 0.0


@pytest.mark.parametrize("groups", [1, 2])
def test_strided_adjoints_in_matryoshka_list(groups):
    """Backward of the DCVNet cost-volume list (fine scale at stride 4 through the strided adjoints, any group count) against
    the oracle, with the fused LeakyReLU mask."""
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    rng = np.random.default_rng(31)
    x1 = [rng.standard_normal((2, 8, 32, 64)).astype(np.float32), rng.standard_normal((2, 16, 8, 16)).astype(np.float32)]
    x2 = [rng.standard_normal((2, 8, 32, 64)).astype(np.float32), rng.standard_normal((2, 16, 8, 16)).astype(np.float32)]
    m = E.MatryoshkaDilatedCostVolumeList(num_groups=groups, max_displacement=4, encoder_output_strides=[2, 8], dilations=[[1], [1, 2]],
                                          normalize_feat_l2=False, use_relu=False).to(dev())
    t1 = [to_gpu(a, True) for a in x1]
    t2 = [to_gpu(a, True) for a in x2]
    out = m(t1, t2)
    dout = rng.standard_normal(tuple(out.shape)).astype(np.float32)
    out.backward(to_gpu(dout))
    # fine scale: channels [0, groups) of the concatenated volume
    P = 9
    g = dout[:, :groups].reshape(2 * groups, P, P, 8, 16)
    a = x1[0].reshape(2 * groups, 8 // groups, 32, 64)
    b = x2[0].reshape(2 * groups, 8 // groups, 32, 64)
    d1, d2 = O.iter_spatial_correlation_sample_bwd(a, b, g, patch_size=P, stride=4)
    assert_close(t1[0].grad, d1.reshape(2, 8, 32, 64), 5e-5, f"fine-scale dx1 groups={groups}")
    assert_close(t2[0].grad, d2.reshape(2, 8, 32, 64), 5e-5, f"fine-scale dx2 groups={groups}")
