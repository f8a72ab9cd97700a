"""Pins oracle/similarity_oracle.py against vectors produced by the unmodified
reference (tools/make_golden.py).  CPU-only."""

import numpy as np
import pytest

from oracle import similarity_oracle as O


def close(a, b, rtol=2e-5, atol=None):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    assert a.shape == b.shape, (a.shape, b.shape)
    scale = max(np.abs(b).max(), 1e-30)
    err = np.abs(a - b).max()
    tol = rtol * scale if atol is None else atol
    assert err <= tol, f"max abs err {err:.3e} > {tol:.3e} (scale {scale:.3e})"


PAIRWISE = ["pairwise_l4_r4", "pairwise_l3_r3", "pairwise_scaled"]


@pytest.mark.parametrize("name", PAIRWISE)
def test_pairwise_forward(golden, name):
    g = golden(name)
    L, r = int(g["levels"]), int(g["radius"])
    pyr = O.corr_pyramid(g["f1"], g["f2"], L)
    if "corr" in g:
        close(O.corr_all_pairs(g["f1"], g["f2"]), g["corr"])
    for i in range(L):
        close(pyr[i], g[f"pyr{i}"])
    # lookup on the *reference's* pyramid isolates K3
    out = O.corr_lookup([g[f"pyr{i}"] for i in range(L)], g["coords"], r)
    close(out, g["out"])
    assert out.dtype == np.float32 and out.flags["C_CONTIGUOUS"]


@pytest.mark.parametrize("name", PAIRWISE)
def test_pairwise_backward(golden, name):
    g = golden(name)
    L, r = int(g["levels"]), int(g["radius"])
    shapes = [g[f"pyr{i}"].shape[-2:] for i in range(L)]
    dpyr = O.corr_lookup_bwd(shapes, g["coords"], g["dout"], r)
    # the reference's .grad of level i also carries the pool-adjoint of level i+1
    total = dpyr[-1]
    close(total, g[f"dpyr{L - 1}"], rtol=5e-5)
    for i in range(L - 2, -1, -1):
        total = dpyr[i] + O.avg_pool2x2_bwd(total, shapes[i])
        close(total, g[f"dpyr{i}"], rtol=5e-5)
    df1, df2 = O.corr_pyramid_bwd(g["f1"], g["f2"], dpyr)
    close(df1, g["df1"], rtol=5e-5)
    close(df2, g["df2"], rtol=5e-5)


@pytest.mark.parametrize("name", ["lookup_dcoords_l4_r4", "lookup_dcoords_l2_r1"])
def test_lookup_coords_gradient(golden, name):
    # the reference's autograd through coords/2**l, the normalisation and grid_sample's grid argument
    g = golden(name)
    L, r = int(g["levels"]), int(g["radius"])
    pyr = O.corr_pyramid(g["f1"], g["f2"], L)
    close(O.corr_lookup(pyr, g["coords"], r), g["out"])
    close(O.corr_lookup_bwd_coords(pyr, g["coords"], g["dout"], r), g["dcoords"], rtol=5e-5)


SAMPLER = ["sampler_p9", "sampler_flownetc", "sampler_stride2", "sampler_aniso", "sampler_dcv_s4", "sampler_dil16"]


def _kw(g):
    kw = {}
    for k, v in g.items():
        if k.startswith("p_"):
            kw[k[2:]] = int(v) if v.ndim == 0 else tuple(int(t) for t in v)
    return kw


@pytest.mark.parametrize("name", SAMPLER)
def test_sampler(golden, name):
    g = golden(name)
    kw = _kw(g)
    close(O.iter_spatial_correlation_sample(g["x1"], g["x2"], **kw), g["out"])
    kwb = {k: v for k, v in kw.items() if k in ("patch_size", "stride", "padding", "dilation_patch")}
    d1, d2 = O.iter_spatial_correlation_sample_bwd(g["x1"], g["x2"], g["dout"], **kwb)
    close(d1, g["dx1"], rtol=5e-5)
    close(d2, g["dx2"], rtol=5e-5)


def test_sampler_rejects():
    x = np.zeros((1, 2, 4, 4), np.float32)
    for kw in (dict(kernel_size=3), dict(dilation=2), dict(patch_size=4)):
        with pytest.raises(NotImplementedError):
            O.iter_spatial_correlation_sample(x, x, **kw)


def test_corr_layer(golden):
    g = golden("corr_layer")
    close(O.correlation_layer(g["x1"], g["x2"]), g["out"])
    close(O.correlation_layer(g["x1"], g["x2"], 3, 3), g["out_p2_d3"])
    # CorrelationLayer == K4(P=9)/C reshaped
    k4 = O.iter_spatial_correlation_sample(g["x1"], g["x2"], patch_size=9)
    b, p, _, h, w = k4.shape
    close(k4.reshape(b, p * p, h, w) / g["x1"].shape[1], g["out"])


def test_matryoshka(golden):
    g = golden("matryoshka")
    close(O.matryoshka_cost_volume(g["x1"], g["x2"]), g["out_default"])
    close(O.matryoshka_cost_volume(g["x1"], g["x2"], 4, 2, 2, [1, 3], True), g["out_g4"])
    np.testing.assert_array_equal(O.concentric_offsets([1, 2, 3, 5, 9, 16], 4), g["offsets_default"])
    np.testing.assert_array_equal(O.concentric_offsets([1, 3], 2), g["offsets_g4"])


def test_matryoshka_list(golden):
    g = golden("matryoshka_list")
    xa, xb = [g["xa0"], g["xa1"]], [g["xb0"], g["xb1"]]
    close(O.matryoshka_cost_volume_list(xa, xb), g["out_default"])
    close(O.matryoshka_cost_volume_list(xa, xb, normalize_feat_l2=True, use_relu=True), g["out_norm_relu"])
    np.testing.assert_array_equal(O.matryoshka_offsets_2d(), g["offsets_2d"])


def test_concentric_offsets_reference_table():
    # the only value-level goldens the reference's own tests hold near this path:
    # tests/test_utils.py:86-127 (concentric_offsets / get_flow_offsets tables)
    offs = O.concentric_offsets([1, 2, 3, 5, 9, 16], 4)
    assert offs.shape == (6, 9)
    np.testing.assert_array_equal(offs[0], np.arange(-4, 5))
    np.testing.assert_array_equal(offs[5], np.arange(-4, 5) * 16)
    o2 = O.matryoshka_offsets_2d()
    assert o2.shape == (7, 9, 9, 2)
    assert o2[0, 0, 0].tolist() == [-8.0, -8.0] and o2[6, 8, 0].tolist() == [512.0, -512.0]


def test_warp(golden):
    g = golden("warp")
    close(O.warp(g["x"], g["flow"]), g["out"])


def test_pool_floor_and_adjoint():
    rng = np.random.default_rng(0)
    x = rng.standard_normal((3, 1, 7, 9)).astype(np.float32)
    y = O.avg_pool2x2(x)
    assert y.shape == (3, 1, 3, 4)
    dy = rng.standard_normal(y.shape).astype(np.float32)
    dx = O.avg_pool2x2_bwd(dy, (7, 9))
    assert np.all(dx[..., 6, :] == 0) and np.all(dx[..., :, 8] == 0)
    np.testing.assert_allclose((y * dy).sum(), (x * dx).sum(), rtol=1e-5)


@pytest.mark.parametrize("name", ["convex_up_s8", "convex_up_s4"])
def test_convex_upsample(golden, name):
    g = golden(name)
    s = int(g["out_stride"])
    close(O.convex_upsample_flow(g["flow"], g["logits"], s), g["out"])
    dflow, dlogits = O.convex_upsample_flow_bwd(g["flow"], g["logits"], s, g["dout"])
    close(dflow, g["dflow"], rtol=5e-5)
    close(dlogits, g["dlogits"], rtol=5e-5)


@pytest.mark.parametrize("name", ["vcn_corr_md3", "vcn_corr_md4_f2"])
def test_vcn_corr_oracle_vs_reference_golden(golden, name):
    g = golden(name)
    out = O.vcn_corr(g["f1"], g["f2"], int(g["max_disp"]), int(g["factorization"]))
    assert out.shape == g["out"].shape
    np.testing.assert_allclose(out, g["out"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("name", ["dicl_volume_u3v3", "dicl_volume_u2v1_keep"])
def test_dicl_volume_oracle_vs_reference_golden(golden, name):
    g = golden(name)
    out = O.dicl_cost_volume(g["x"], g["y"], int(g["max_u"]), int(g["max_v"]), bool(g["remove_warp_hole"]))
    assert out.shape == g["out"].shape
    np.testing.assert_array_equal(out, g["out"])  # pure data movement: bit-exact
