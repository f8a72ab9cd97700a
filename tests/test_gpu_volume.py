"""GPU parity, SURVEY.md §8f-4: VCN's per-channel correlation (reference ezflow/models/vcn.py:162-206), DICL's
displacement-stacked volume (ezflow/similarity/learnable_cost.py:181-226) and the on-demand (volume-free) RAFT lookup
(pairwise.py:43-69), against golden vectors produced by the unmodified reference and against the numpy oracle.
Data movement is bit-exact; the fp32 products / dot products are compared at 1e-6 / 2e-5."""

import numpy as np
import pytest
import torch

from gpu_util import assert_close, dev, to_gpu
from oracle import similarity_oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["vcn_corr_md3", "vcn_corr_md4_f2"])
def test_vcn_corr_vs_reference_golden(golden, name):
    import ezflow_b200 as E

    g = golden(name)
    f1, f2 = to_gpu(g["f1"], True), to_gpu(g["f2"], True)
    out = E.vcn_corr(f1, f2, int(g["max_disp"]), int(g["factorization"]))
    assert out.is_contiguous()
    assert_close(out, g["out"], 1e-6, "vcn corr")
    (out * to_gpu(g["dout"])).sum().backward()
    assert_close(f1.grad, g["df1"], 2e-5, "vcn corr dfeatures1")
    assert_close(f2.grad, g["df2"], 2e-5, "vcn corr dfeatures2")


def test_vcn_corr_shapes_vs_oracle():
    import ezflow_b200 as E

    rng = np.random.default_rng(3)
    for (b, c, h, w, md, fac) in [(1, 3, 5, 7, 6, 1), (2, 16, 33, 40, 4, 2), (1, 1, 1, 1, 2, 1)]:  # window wider than the map; ragged; 1x1
        f1 = rng.standard_normal((b, c, h, w)).astype(np.float32)
        f2 = rng.standard_normal((b, c, h, w)).astype(np.float32)
        with torch.no_grad():
            out = E.vcn_corr(to_gpu(f1), to_gpu(f2), md, fac)
        assert_close(out, O.vcn_corr(f1, f2, md, fac), 1e-6, f"vcn corr {(b, c, h, w, md, fac)}")


@pytest.mark.parametrize("name", ["dicl_volume_u3v3", "dicl_volume_u2v1_keep"])
def test_dicl_volume_vs_reference_golden(golden, name):
    import ezflow_b200 as E

    g = golden(name)
    x, y = to_gpu(g["x"], True), to_gpu(g["y"], True)
    vol = E.dicl_cost_volume(x, y, int(g["max_u"]), int(g["max_v"]), bool(g["remove_warp_hole"]))
    assert vol.is_contiguous() and tuple(vol.shape) == g["out"].shape
    np.testing.assert_array_equal(vol.detach().cpu().numpy(), g["out"])  # pure data movement: bit-exact
    (vol * to_gpu(g["dout"])).sum().backward()
    assert_close(x.grad, g["dx"], 2e-6, "dicl volume dx")
    assert_close(y.grad, g["dy"], 2e-6, "dicl volume dy")


def test_dicl_volume_shapes_vs_oracle():
    import ezflow_b200 as E

    rng = np.random.default_rng(4)
    for (b, c, h, w, mu, mv, holes) in [(1, 32, 24, 32, 3, 3, True), (2, 3, 4, 5, 6, 2, True), (1, 8, 7, 9, 0, 0, False)]:
        x = rng.standard_normal((b, c, h, w)).astype(np.float32)
        y = rng.standard_normal((b, c, h, w)).astype(np.float32)
        y[:, :, : h // 2, : w // 3] = 0.0
        with torch.no_grad():
            vol = E.dicl_cost_volume(to_gpu(x), to_gpu(y), mu, mv, holes)
        np.testing.assert_array_equal(vol.cpu().numpy(), O.dicl_cost_volume(x, y, mu, mv, holes))


@pytest.mark.parametrize("cfg", [(2, 32, 16, 21, 4, 4), (1, 24, 10, 14, 3, 3), (1, 7, 9, 12, 2, 0), (1, 256, 46, 62, 4, 4)])
def test_ondemand_lookup_vs_oracle(cfg):
    """Same output as the pyramid path's lookup, from the feature maps alone (fp32)."""
    import ezflow_b200 as E

    b, c, h, w, L, r = cfg
    rng = np.random.default_rng(c + h)
    f1 = rng.standard_normal((b, c, h, w)).astype(np.float32)
    f2 = rng.standard_normal((b, c, h, w)).astype(np.float32)
    coords = O.coords_grid(b, h, w) + rng.uniform(-1.5 * r - 2, 1.5 * r + 2, (b, 2, h, w)).astype(np.float32)
    ref = O.corr_lookup(O.corr_pyramid(f1, f2, L), coords, r)
    with torch.no_grad():
        fn = E.OnDemandPairwise4DCorr(to_gpu(f1), to_gpu(f2), num_levels=L, corr_radius=r)
        out = fn(to_gpu(coords))
        assert tuple(out.shape) == ref.shape and out.is_contiguous()
        assert_close(out, ref, 2e-5, f"on-demand lookup {cfg}")
        # and against the pyramid path of this repo (fp16 tensor-core volume): the north_star's 1e-3
        pyr = E.MutliScalePairwise4DCorr(to_gpu(f1), to_gpu(f2), num_levels=L, corr_radius=r)(to_gpu(coords))
        assert_close(out, pyr, 1e-3, "on-demand vs pyramid path")


def test_ondemand_edge_cases():
    import ezflow_b200 as E

    z = torch.zeros(0, 16, 12, 20, device=dev())
    fn = E.OnDemandPairwise4DCorr(z, z, num_levels=2, corr_radius=3)
    assert tuple(fn(torch.zeros(0, 2, 12, 20, device=dev())).shape) == (0, 98, 12, 20)
    with pytest.raises(NotImplementedError):
        a = torch.zeros(1, 8, 64, 64, device=dev())
        E.OnDemandPairwise4DCorr(a, a, num_levels=5, corr_radius=1)
    with pytest.raises(RuntimeError):
        E.OnDemandPairwise4DCorr(torch.zeros(1, 4, 8, 8), torch.zeros(1, 4, 8, 8))  # CPU tensors: no fallback


@pytest.mark.parametrize("cfg", [(2, 64, 30, 44, 4, 4, 0.4), (1, 40, 17, 23, 3, 3, 1.5), (1, 256, 46, 62, 4, 4, 0.2), (1, 32, 20, 36, 2, 4, 30.0)])
def test_ondemand_tiled_lookup_coherent_and_scattered(cfg):
    """corr_ondemand_tiled_kernel: a smooth flow (the windows of a 4 x 16 query block fit the shared-memory region at every
    level), mild noise (mixed), and wide scatter (every block on the per-pixel fallback), against the oracle and against
    the untiled kernel (EZB_ONDEMAND_UNTILED=1); channel counts off the 32-channel chunk, ragged query blocks."""
    import os

    import ezflow_b200 as E

    b, c, h, w, L, r, noise = cfg
    rng = np.random.default_rng(int(noise * 10) + c)
    f1 = rng.standard_normal((b, c, h, w)).astype(np.float32)
    f2 = rng.standard_normal((b, c, h, w)).astype(np.float32)
    yy, xx = np.meshgrid(np.linspace(0, 3, h), np.linspace(0, 3, w), indexing="ij")
    flow = np.stack([6 * np.sin(xx + yy), 5 * np.cos(xx - yy)]).astype(np.float32)[None]
    coords = O.coords_grid(b, h, w) + flow + rng.normal(0, noise, (b, 2, h, w)).astype(np.float32)
    ref = O.corr_lookup(O.corr_pyramid(f1, f2, L), coords, r)
    with torch.no_grad():
        fn = E.OnDemandPairwise4DCorr(to_gpu(f1), to_gpu(f2), num_levels=L, corr_radius=r)
        out = fn(to_gpu(coords))
        os.environ["EZB_ONDEMAND_UNTILED"] = "1"
        try:
            old = fn(to_gpu(coords))
        finally:
            del os.environ["EZB_ONDEMAND_UNTILED"]
    assert_close(out, ref, 2e-5, f"tiled on-demand lookup {cfg}")
    assert_close(out, old, 2e-5, "tiled vs untiled kernel")
