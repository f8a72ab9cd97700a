"""GPU parity: RAFT all-pairs correlation pyramid + lookup (K1, K2, K3, K3b, K1b) through the drop-in
class, against (a) golden vectors produced by the unmodified reference and (b) the numpy oracle.

Tolerances (north_star): correlation volumes <= 1e-3 relative error (fp16 tensor-core contraction vs the
reference's fp32); lookups inherit that bound."""

import os

import numpy as np
import pytest
import torch

from gpu_util import assert_close, dev, rel_fro, to_gpu

pytestmark = pytest.mark.gpu

VOL_TOL = 1e-3   # correlation volumes / pyramid levels (north_star)
LOOK_TOL = 1e-3  # lookup output, relative to the reference's output
GRAD_TOL = 2e-3  # adjoints: two chained fp32 reductions in different orders + fp16 pyramid not involved


@pytest.fixture(autouse=True)
def _restore_dtype():
    import ezflow_b200 as E

    old = E.get_pyramid_dtype()
    yield
    E.set_pyramid_dtype(old)


PAIRWISE = ["pairwise_l4_r4", "pairwise_l3_r3", "pairwise_scaled"]


@pytest.mark.parametrize("store", ["fp16", "fp32"])
@pytest.mark.parametrize("name", PAIRWISE)
def test_pyramid_and_lookup_vs_reference_golden(golden, name, store):
    import ezflow_b200 as E

    g = golden(name)
    L, r = int(g["levels"]), int(g["radius"])
    E.set_pyramid_dtype(store)
    obj = E.MutliScalePairwise4DCorr(to_gpu(g["f1"]), to_gpu(g["f2"]), num_levels=L, corr_radius=r)
    assert obj.num_levels == L and obj.corr_radius == r
    pyr = obj.corr_pyramid
    assert len(pyr) == L
    for i in range(L):
        assert tuple(pyr[i].shape) == g[f"pyr{i}"].shape
        assert_close(pyr[i], g[f"pyr{i}"], VOL_TOL, f"{name} level {i} ({store})")
    out = obj(to_gpu(g["coords"]))
    assert out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == g["out"].shape
    assert_close(out, g["out"], LOOK_TOL, f"{name} lookup ({store})")


@pytest.mark.parametrize("name", ["pairwise_l3_r3", "pairwise_scaled"])
def test_corr_static(golden, name):
    import ezflow_b200 as E

    g = golden(name)
    corr = E.MutliScalePairwise4DCorr.corr(to_gpu(g["f1"]), to_gpu(g["f2"]))
    assert tuple(corr.shape) == g["corr"].shape
    assert_close(corr, g["corr"], VOL_TOL, f"{name} corr()")


@pytest.mark.parametrize("name", PAIRWISE)
def test_backward_vs_reference_golden(golden, name):
    import ezflow_b200 as E

    g = golden(name)
    L, r = int(g["levels"]), int(g["radius"])
    f1, f2 = to_gpu(g["f1"], True), to_gpu(g["f2"], True)
    obj = E.MutliScalePairwise4DCorr(f1, f2, num_levels=L, corr_radius=r)
    out = obj(to_gpu(g["coords"]))
    (out * to_gpu(g["dout"])).sum().backward()
    assert_close(f1.grad, g["df1"], GRAD_TOL, f"{name} dfmap1")
    assert_close(f2.grad, g["df2"], GRAD_TOL, f"{name} dfmap2")


@pytest.mark.parametrize("store", ["fp16", "fp32"])
@pytest.mark.parametrize("name", ["lookup_dcoords_l4_r4", "lookup_dcoords_l2_r1"])
def test_coords_gradient_vs_reference_golden(golden, name, store):
    """RAFT detaches coords (models/raft.py:115), but the reference's lookup is differentiable in them: so is this one."""
    import ezflow_b200 as E

    g = golden(name)
    L, r = int(g["levels"]), int(g["radius"])
    E.set_pyramid_dtype(store)
    coords = to_gpu(g["coords"], True)
    obj = E.MutliScalePairwise4DCorr(to_gpu(g["f1"]), to_gpu(g["f2"]), num_levels=L, corr_radius=r)  # fmaps without grad
    out = obj(coords)
    assert_close(out, g["out"], LOOK_TOL, f"{name} lookup ({store})")
    (out * to_gpu(g["dout"])).sum().backward()
    assert tuple(coords.grad.shape) == g["dcoords"].shape
    assert_close(coords.grad, g["dcoords"], GRAD_TOL, f"{name} dcoords ({store})")


def test_coords_and_feature_gradients_together():
    """coords = coords0 + flow(f1): both adjoints in one backward pass, against the oracle's chain."""
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    rng = np.random.default_rng(7)
    B, C, H, W, L, r = 2, 32, 12, 18, 3, 2
    f1 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    f2 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    coords = O.coords_grid(B, H, W) + rng.uniform(-4, 4, (B, 2, H, W)).astype(np.float32)
    dout = rng.standard_normal((B, L * (2 * r + 1) ** 2, H, W)).astype(np.float32)
    t1, t2, tc = to_gpu(f1, True), to_gpu(f2, True), to_gpu(coords, True)
    out = E.MutliScalePairwise4DCorr(t1, t2, num_levels=L, corr_radius=r)(tc * 1.0)  # non-leaf coords
    (out * to_gpu(dout)).sum().backward()
    pyr = O.corr_pyramid(f1, f2, L)
    dpyr = O.corr_lookup_bwd([p.shape[-2:] for p in pyr], coords, dout, r)
    d1, d2 = O.corr_pyramid_bwd(f1, f2, dpyr)
    assert_close(tc.grad, O.corr_lookup_bwd_coords(pyr, coords, dout, r), GRAD_TOL, "dcoords")
    assert_close(t1.grad, d1, GRAD_TOL, "dfmap1")
    assert_close(t2.grad, d2, GRAD_TOL, "dfmap2")


@pytest.mark.parametrize("shape", [(2, 40, 9, 11), (2, 256, 12, 20), (1, 256, 11, 13)])
def test_backward_vs_oracle_shapes(shape):
    """Adjoints vs the numpy oracle on shapes the golden fixtures do not hold: H*W % 4 != 0 (TMA cannot describe the
    operands: fold + fp32 CUDA-core GEMMs) and H*W % 4 == 0 with C = 256 (tf32 tcgen05 GEMMs over all levels), B = 2."""
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    rng = np.random.default_rng(21)
    B, C, H, W = shape
    L, r = 3, 2
    f1n = rng.standard_normal(shape).astype(np.float32)
    f2n = rng.standard_normal(shape).astype(np.float32)
    coords = O.coords_grid(B, H, W) + rng.uniform(-3, 3, (B, 2, H, W)).astype(np.float32)
    dout = rng.standard_normal((B, L * (2 * r + 1) ** 2, H, W)).astype(np.float32)
    shapes = [(H >> l, W >> l) for l in range(L)]
    dpyr = O.corr_lookup_bwd(shapes, coords, dout, r)
    df1, df2 = O.corr_pyramid_bwd(f1n, f2n, dpyr)
    f1, f2 = to_gpu(f1n, True), to_gpu(f2n, True)
    obj = E.MutliScalePairwise4DCorr(f1, f2, num_levels=L, corr_radius=r)
    (obj(to_gpu(coords)) * to_gpu(dout)).sum().backward()
    assert_close(f1.grad, df1, GRAD_TOL, f"{shape} dfmap1")
    assert_close(f2.grad, df2, GRAD_TOL, f"{shape} dfmap2")


def test_backward_two_lookups_accumulate(golden):
    """Two lookups through one pyramid: gradients add (RAFT calls the object 12x, models/raft.py:116)."""
    import ezflow_b200 as E

    g = golden("pairwise_l3_r3")
    L, r = int(g["levels"]), int(g["radius"])
    f1, f2 = to_gpu(g["f1"], True), to_gpu(g["f2"], True)
    obj = E.MutliScalePairwise4DCorr(f1, f2, num_levels=L, corr_radius=r)
    c = to_gpu(g["coords"])
    d = to_gpu(g["dout"])
    ((obj(c) * d).sum() + (obj(c) * d).sum()).backward()
    assert_close(f1.grad, 2 * g["df1"], GRAD_TOL, "2x dfmap1")
    assert_close(f2.grad, 2 * g["df2"], GRAD_TOL, "2x dfmap2")


@pytest.mark.parametrize("store", ["fp16", "fp32"])
def test_cfg1_shape_vs_oracle(store):
    """BASELINE config 1 feature size (368x496 -> 46x62, C=256): CUDA path vs numpy oracle."""
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    rng = np.random.default_rng(1234)
    f1 = rng.standard_normal((1, 256, 46, 62)).astype(np.float32)
    f2 = rng.standard_normal((1, 256, 46, 62)).astype(np.float32)
    coords = O.coords_grid(1, 46, 62) + rng.uniform(-8, 8, (1, 2, 46, 62)).astype(np.float32)
    ref_pyr = O.corr_pyramid(f1, f2, 4)
    ref_out = O.corr_lookup(ref_pyr, coords, 4)
    E.set_pyramid_dtype(store)
    obj = E.MutliScalePairwise4DCorr(to_gpu(f1), to_gpu(f2), num_levels=4, corr_radius=4)
    pyr = obj.corr_pyramid
    for i in range(4):
        rf, mr = assert_close(pyr[i], ref_pyr[i], VOL_TOL, f"cfg1 level {i} ({store})")
    out = obj(to_gpu(coords))
    assert tuple(out.shape) == (1, 324, 46, 62)
    assert_close(out, ref_out, LOOK_TOL, f"cfg1 lookup ({store})")


def test_batch_and_odd_channels_vs_oracle():
    """B=3, C=96 (Cp padded to 128), odd H/W so every level drops a row/col; radius 3, 3 levels."""
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    rng = np.random.default_rng(7)
    f1 = rng.standard_normal((3, 96, 19, 37)).astype(np.float32)
    f2 = rng.standard_normal((3, 96, 19, 37)).astype(np.float32)
    f1[1] *= 40.0  # per-pair scales differ by orders of magnitude
    f2[2] *= 1e-3
    coords = O.coords_grid(3, 19, 37) + rng.uniform(-5, 5, (3, 2, 19, 37)).astype(np.float32)
    ref_pyr = O.corr_pyramid(f1, f2, 3)
    ref_out = O.corr_lookup(ref_pyr, coords, 3)
    obj = E.MutliScalePairwise4DCorr(to_gpu(f1), to_gpu(f2), num_levels=3, corr_radius=3)
    n = 19 * 37
    pyr = obj.corr_pyramid
    for b in range(3):  # per pair: scales differ, so relative error is judged per pair
        for i in range(3):
            assert_close(pyr[i][b * n : (b + 1) * n], ref_pyr[i][b * n : (b + 1) * n], VOL_TOL, f"pair {b} level {i}")
        assert_close(obj(to_gpu(coords))[b], ref_out[b], LOOK_TOL, f"pair {b} lookup")


def test_lookup_edge_coordinates():
    """Windows fully / partially outside, exactly-integer coordinates, huge values: zeros padding."""
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    rng = np.random.default_rng(3)
    H, W = 12, 20
    f1 = rng.standard_normal((1, 64, H, W)).astype(np.float32)
    f2 = rng.standard_normal((1, 64, H, W)).astype(np.float32)
    coords = O.coords_grid(1, H, W)
    coords[0, 0, 0, :] = -50.0       # far left
    coords[0, 1, 1, :] = 1e6         # far below
    coords[0, 0, 2, :] = W - 1 + 3.5 # straddling the right border
    coords[0, :, 3, :] = -0.5        # straddling the top-left corner
    coords[0, 0, 4, :] = -1e9
    E.set_pyramid_dtype("fp32")
    obj = E.MutliScalePairwise4DCorr(to_gpu(f1), to_gpu(f2), num_levels=2, corr_radius=4)
    out = obj(to_gpu(coords)).cpu().numpy()
    ref = O.corr_lookup(O.corr_pyramid(f1, f2, 2), coords, 4)
    assert np.isfinite(out).all()
    assert np.all(out[0, :, 0, :] == 0) and np.all(out[0, :, 1, :] == 0) and np.all(out[0, :, 4, :] == 0)
    assert_close(out, ref, LOOK_TOL, "edge lookup")


def test_sharding_invariance():
    """Per-pair scaling => a pair's result does not depend on which batch (GPU shard) it is in: bit-exact."""
    import ezflow_b200 as E

    g = torch.Generator(device="cpu").manual_seed(5)
    f1 = torch.randn(4, 128, 16, 24, generator=g).to(dev())
    f2 = torch.randn(4, 128, 16, 24, generator=g).to(dev())
    f2[3] *= 100
    coords = (E.coords_grid(4, 16, 24, device=dev()) + torch.rand(4, 2, 16, 24, generator=g).to(dev()) * 6 - 3)
    whole = E.MutliScalePairwise4DCorr(f1, f2)(coords)
    for lo, hi in ((0, 2), (2, 4)):
        part = E.MutliScalePairwise4DCorr(f1[lo:hi], f2[lo:hi])(coords[lo:hi])
        assert torch.equal(part, whole[lo:hi])


def test_more_than_four_levels_is_unsupported():
    import ezflow_b200 as E

    x = torch.randn(1, 64, 32, 48, device=dev())
    with pytest.raises(NotImplementedError):
        E.MutliScalePairwise4DCorr(x, x, num_levels=5, corr_radius=1)


def test_errors():
    import ezflow_b200 as E

    with pytest.raises(RuntimeError):
        E.MutliScalePairwise4DCorr(torch.rand(1, 8, 8, 8), torch.rand(1, 8, 8, 8))  # CPU tensors: no fallback
    with pytest.raises(NotImplementedError):
        E.MutliScalePairwise4DCorr(torch.rand(1, 320, 8, 8, device=dev()), torch.rand(1, 320, 8, 8, device=dev()))
    with pytest.raises(ValueError):
        E.MutliScalePairwise4DCorr(torch.rand(1, 8, 4, 4, device=dev()), torch.rand(1, 8, 4, 4, device=dev()), num_levels=4)


def test_full_size_1080p_properties():
    """BASELINE headline size (135x240, C=256): size-independent checks instead of a full oracle run.
    (a) sampled rows of level 0 vs an fp32 matmul; (b) pooled levels are averages of the level below;
    (c) lookup at integer coordinates returns pyramid entries; (d) linearity in fmap1."""
    import ezflow_b200 as E

    H, W, C = 135, 240, 256
    N = H * W
    g = torch.Generator(device="cpu").manual_seed(11)
    f1 = torch.randn(1, C, H, W, generator=g).to(dev())
    f2 = torch.randn(1, C, H, W, generator=g).to(dev())
    obj = E.MutliScalePairwise4DCorr(f1, f2)
    st = obj._state
    qs = torch.tensor([0, 1, 239, 240, 12345, N - 1], device=dev())
    lv = [st.level_tensor(l, queries=qs) for l in range(4)]
    deq = st.deq[0].item()
    # (a)
    ref_rows = (f1.view(C, N)[:, qs].t().double() @ f2.view(C, N).double() / 16.0).view(-1, H, W)
    got_rows = lv[0][:, 0].double() * deq
    assert rel_fro(got_rows.cpu().numpy(), ref_rows.cpu().numpy()) <= VOL_TOL
    # (b) level l+1 == floor-avg-pool of the exact fp32 level below (computed from ref rows)
    cur = ref_rows
    for l in range(1, 4):
        h, w = cur.shape[-2] // 2, cur.shape[-1] // 2
        cur = cur[:, : 2 * h, : 2 * w].reshape(-1, h, 2, w, 2).mean(dim=(2, 4))
        got = lv[l][:, 0].double() * deq
        assert tuple(got.shape[-2:]) == (h, w)
        assert rel_fro(got.cpu().numpy(), cur.cpu().numpy()) <= VOL_TOL
    # (c) integer coords: centre tap (i=j=r) of level 0 equals corr[q, q]
    coords = E.coords_grid(1, H, W, device=dev())
    out = obj(coords)
    centre = out[0, 4 * 9 + 4].reshape(-1)[qs].double()
    diag = torch.stack([got_rows[i].reshape(-1)[q] for i, q in enumerate(qs.tolist())])
    assert torch.allclose(centre, diag, rtol=1e-5, atol=1e-6)
    # (d) linearity: scaling fmap1 by 3 scales every lookup by 3 (power-of-two pre-scale makes this near-exact)
    out3 = E.MutliScalePairwise4DCorr(3.0 * f1, f2)(coords)
    assert rel_fro(out3.cpu().numpy(), (3.0 * out).cpu().numpy()) <= 5e-4


@pytest.mark.parametrize("store", ["fp16", "fp32"])
@pytest.mark.parametrize("cfg", [(1, 8, 4, 4, 1, 1), (1, 16, 5, 7, 2, 2), (2, 3, 2, 9, 1, 4), (1, 256, 3, 3, 1, 4), (1, 1, 16, 16, 4, 4)])
def test_tiny_and_ragged_shapes_vs_oracle(cfg, store):
    """Feature maps smaller than one GEMM tile / one query group, single channel, one pyramid level: CUDA path vs oracle."""
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    B, C, H, W, L, r = cfg
    rng = np.random.default_rng(100 + H * W)
    f1 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    f2 = rng.standard_normal((B, C, H, W)).astype(np.float32)
    coords = O.coords_grid(B, H, W) + rng.uniform(-2, 2, (B, 2, H, W)).astype(np.float32)
    ref = O.corr_lookup(O.corr_pyramid(f1, f2, L), coords, r)
    E.set_pyramid_dtype(store)
    out = E.MutliScalePairwise4DCorr(to_gpu(f1), to_gpu(f2), num_levels=L, corr_radius=r)(to_gpu(coords))
    assert_close(out, ref, LOOK_TOL, f"{cfg} ({store})")


def test_full_size_is_deterministic():
    """Headline size, two pairs: repeated builds + lookups are bit-identical (no races in the TMA / tcgen05 pipelines)."""
    import ezflow_b200 as E

    g = torch.Generator(device="cpu").manual_seed(5)
    f1 = torch.randn(2, 256, 135, 240, generator=g).to(dev())
    f2 = torch.randn(2, 256, 135, 240, generator=g).to(dev())
    c = E.coords_grid(2, 135, 240, device=dev()) + torch.randn(2, 2, 135, 240, generator=g).to(dev()) * 5
    ref = E.MutliScalePairwise4DCorr(f1, f2)(c)
    for _ in range(3):
        assert torch.equal(E.MutliScalePairwise4DCorr(f1, f2)(c), ref)


def test_inplace_edit_between_forward_and_backward_raises():
    """The feature maps are saved through autograd (save_for_backward), so modifying them in place before backward()
    trips PyTorch's version check exactly as it does for the reference's matmul."""
    import ezflow_b200 as E

    rng = np.random.default_rng(2)
    f1 = to_gpu(rng.standard_normal((1, 16, 8, 12)).astype(np.float32), True)
    f2 = to_gpu(rng.standard_normal((1, 16, 8, 12)).astype(np.float32), True)
    a, b = f1 * 1.0, f2 * 1.0  # non-leaf tensors that may be edited in place
    fn = E.MutliScalePairwise4DCorr(a, b, num_levels=2, corr_radius=2)
    out = fn(E.coords_grid(1, 8, 12).to(dev()))
    a.add_(1.0)
    with pytest.raises(RuntimeError, match="modified by an inplace operation"):
        out.sum().backward()


def test_aborted_backward_leaves_no_stale_gradient():
    """A backward pass that raises after some lookup adjoints already accumulated must not leak into the next pass."""
    import ezflow_b200 as E

    rng = np.random.default_rng(3)
    x1 = rng.standard_normal((1, 16, 8, 12)).astype(np.float32)
    x2 = rng.standard_normal((1, 16, 8, 12)).astype(np.float32)
    coords = E.coords_grid(1, 8, 12).to(dev()) + 0.3

    class Boom(torch.autograd.Function):
        @staticmethod
        def forward(ctx, t):
            return t.clone()

        @staticmethod
        def backward(ctx, g):
            raise ValueError("boom")

    def grads(with_abort):
        f1, f2 = to_gpu(x1, True), to_gpu(x2, True)
        fn = E.MutliScalePairwise4DCorr(f1, f2, num_levels=2, corr_radius=2)
        if with_abort:
            # the engine runs later-created nodes first: the lookup adjoint fills the accumulator, then Boom (created before
            # the lookup) aborts the pass before the pyramid adjoint has consumed it
            boom = Boom.apply(fn._token * 1.0)
            out = fn(coords)
            with pytest.raises(ValueError):
                (out.sum() + boom).backward()
            assert fn._state.pending  # the aborted pass did leave its deferred lookup adjoints behind
            f1.grad = f2.grad = None
        fn(coords).sum().backward()
        return f1.grad.clone(), f2.grad.clone()

    g1, g2 = grads(False)
    h1, h2 = grads(True)
    assert_close(h1, g1, 1e-6, "df1 after an aborted pass")
    assert_close(h2, g2, 1e-6, "df2 after an aborted pass")


def _train_like_step(E, f1n, f2n, coords_list, douts, L, r):
    f1, f2 = to_gpu(f1n, True), to_gpu(f2n, True)
    fn = E.MutliScalePairwise4DCorr(f1, f2, num_levels=L, corr_radius=r)
    loss = 0
    for c, d in zip(coords_list, douts):
        loss = loss + (fn(to_gpu(c)) * to_gpu(d)).sum()
    loss.backward()
    return f1.grad.clone(), f2.grad.clone()


def test_ranged_adjoint_matches_dense_and_oracle():
    """The adjoint restricted to the support of the lookup windows (ezb_corr_grad_ranges + ezb_corr_pyramid_bwd_ex +
    ezb_corr_grad_rezero) against the dense adjoint (EZB_BWD_DENSE=1) and the oracle: several lookups whose windows wander
    (incl. partly and wholly outside the levels), H*W spanning several 256-query / 256-target tiles, two pairs."""
    import ezflow_b200 as E
    from ezflow_b200.similarity import pairwise as PW
    from oracle import similarity_oracle as O

    rng = np.random.default_rng(77)
    B, C, H, W, L, r = 2, 48, 28, 36, 3, 3  # N = 1008: four 256-query tiles, the last one ragged
    f1n = rng.standard_normal((B, C, H, W)).astype(np.float32)
    f2n = rng.standard_normal((B, C, H, W)).astype(np.float32)
    grid = O.coords_grid(B, H, W)
    coords_list = [grid + rng.normal(0, 1.5 + 2 * k, (B, 2, H, W)).astype(np.float32) for k in range(4)]
    coords_list[3][1, :, :5, :] += 100.0  # some windows entirely outside every level
    douts = [rng.standard_normal((B, L * (2 * r + 1) ** 2, H, W)).astype(np.float32) for _ in coords_list]
    shapes = [(H >> l, W >> l) for l in range(L)]
    dpyr = None
    for c, d in zip(coords_list, douts):
        g = O.corr_lookup_bwd(shapes, c, d, r)
        dpyr = g if dpyr is None else [a + b for a, b in zip(dpyr, g)]
    d1, d2 = O.corr_pyramid_bwd(f1n, f2n, dpyr)

    E.release_gradient_pool()
    r1, r2 = _train_like_step(E, f1n, f2n, coords_list, douts, L, r)
    assert len(PW._GRAD_POOL) == 1, "the ranged path should have handed its gradient pyramid to the pool"
    for lv in next(iter(PW._GRAD_POOL.values())):
        assert float(lv.abs().max()) == 0.0, "the pooled gradient pyramid must be all-zero again"
    os.environ["EZB_BWD_DENSE"] = "1"
    try:
        e1, e2 = _train_like_step(E, f1n, f2n, coords_list, douts, L, r)
    finally:
        del os.environ["EZB_BWD_DENSE"]
    assert_close(r1, d1, GRAD_TOL, "ranged dfmap1 vs oracle")
    assert_close(r2, d2, GRAD_TOL, "ranged dfmap2 vs oracle")
    assert_close(r1, e1, 1e-6, "ranged vs dense dfmap1")  # same products in the same order, only all-zero blocks skipped
    assert_close(r2, e2, 1e-6, "ranged vs dense dfmap2")


def test_ranged_adjoint_pool_reuse_across_steps():
    """Step 2 reuses step 1's gradient pyramid from the pool: different windows, no residue; a pass whose lookups all fall
    outside the levels yields exact zeros."""
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    rng = np.random.default_rng(5)
    B, C, H, W, L, r = 1, 32, 24, 32, 2, 4
    f1n = rng.standard_normal((B, C, H, W)).astype(np.float32)
    f2n = rng.standard_normal((B, C, H, W)).astype(np.float32)
    grid = O.coords_grid(B, H, W)
    shapes = [(H >> l, W >> l) for l in range(L)]
    E.release_gradient_pool()
    for step, shift in enumerate([(-6.3, 2.2), (7.1, -4.6), (0.4, 0.2)]):
        c = grid.copy()
        c[:, 0] += shift[0]
        c[:, 1] += shift[1]
        d = rng.standard_normal((B, L * (2 * r + 1) ** 2, H, W)).astype(np.float32)
        g1, g2 = _train_like_step(E, f1n, f2n, [c], [d], L, r)
        d1, d2 = O.corr_pyramid_bwd(f1n, f2n, O.corr_lookup_bwd(shapes, c, d, r))
        assert_close(g1, d1, GRAD_TOL, f"step {step} dfmap1")
        assert_close(g2, d2, GRAD_TOL, f"step {step} dfmap2")
    far = grid + 1000.0
    d = rng.standard_normal((B, L * (2 * r + 1) ** 2, H, W)).astype(np.float32)
    g1, g2 = _train_like_step(E, f1n, f2n, [far], [d], L, r)
    assert float(g1.abs().max()) == 0.0 and float(g2.abs().max()) == 0.0
    E.release_gradient_pool()


@pytest.mark.parametrize("spread", [0.3, 2.0, 9.0])
def test_merged_lookup_adjoint_matches_per_lookup_launches(spread):
    """ezb_corr_lookup_bwd_multi (all lookups of the pass in one launch, overlapping windows summed in shared memory) against
    one ezb_corr_lookup_bwd launch per lookup (EZB_LOOKUP_BWD_SEPARATE=1) and the oracle: windows that stay together
    (tile path), drift apart (mixed) and scatter (per-lookup path inside the kernel), ragged query blocks, levels whose row
    pitch is not a multiple of 4 (scalar flush)."""
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    rng = np.random.default_rng(int(spread * 10))
    B, C, H, W, L, r = 2, 24, 22, 30, 3, 4  # N = 660 (not a multiple of 16 x ...: 41.25 blocks), w = 30, 15, 7
    f1n = rng.standard_normal((B, C, H, W)).astype(np.float32)
    f2n = rng.standard_normal((B, C, H, W)).astype(np.float32)
    base = O.coords_grid(B, H, W) + rng.uniform(-6, 6, (B, 2, 1, 1)).astype(np.float32)
    coords_list = [base + rng.normal(0, spread, (B, 2, H, W)).astype(np.float32) for _ in range(5)]
    coords_list[2][0, :, 3:6, :] -= 40.0  # windows outside on one side
    douts = [rng.standard_normal((B, L * (2 * r + 1) ** 2, H, W)).astype(np.float32) for _ in coords_list]
    shapes = [(H >> l, W >> l) for l in range(L)]
    dpyr = None
    for c, d in zip(coords_list, douts):
        g = O.corr_lookup_bwd(shapes, c, d, r)
        dpyr = g if dpyr is None else [a + b for a, b in zip(dpyr, g)]
    d1, d2 = O.corr_pyramid_bwd(f1n, f2n, dpyr)
    m1, m2 = _train_like_step(E, f1n, f2n, coords_list, douts, L, r)
    os.environ["EZB_LOOKUP_BWD_SEPARATE"] = "1"
    try:
        s1, s2 = _train_like_step(E, f1n, f2n, coords_list, douts, L, r)
    finally:
        del os.environ["EZB_LOOKUP_BWD_SEPARATE"]
    assert_close(m1, d1, GRAD_TOL, f"merged dfmap1 vs oracle (spread {spread})")
    assert_close(m2, d2, GRAD_TOL, f"merged dfmap2 vs oracle (spread {spread})")
    # same terms in another summation order, then tf32 GEMMs: the truncation can fall differently (measured 9e-5 max)
    assert_close(m1, s1, 3e-4, "merged vs per-lookup dfmap1")
    assert_close(m2, s2, 3e-4, "merged vs per-lookup dfmap2")
    E.release_gradient_pool()


@pytest.mark.parametrize("cfg", [(2, 22, 30, 3, 4, 5), (1, 9, 12, 2, 2, 3), (1, 28, 36, 4, 3, 17)])
def test_grad_ranges_cover_the_lookup_adjoint_support(cfg):
    """C ABI level: whatever ezb_corr_lookup_bwd_multi writes into an all-zero gradient pyramid lies inside the ranges of
    ezb_corr_grad_ranges — ezb_corr_grad_rezero brings the pyramid back to all-zero (the invariant the pooled buffers rely
    on); 17 lookups exercise the 16-lookup chunking, shifted rows put windows wholly outside the levels."""
    import ctypes

    import ezflow_b200 as E
    from ezflow_b200 import _cabi
    from ezflow_b200.similarity.pairwise import grad_layout

    B, H, W, L, r, n = cfg
    K = 2 * r + 1
    g = torch.Generator(device="cpu").manual_seed(n)
    lib, st = _cabi.lib(), _cabi.stream_ptr(dev())
    lv = [torch.zeros(B * H * W * qs, device=dev()) for (_, _, _, qs) in grad_layout(H, W, L)]
    grid = E.coords_grid(B, H, W)
    coords = [(grid + torch.randn(B, 2, H, W, generator=g) * (0.5 + 3 * (k % 3))).to(dev()).contiguous() for k in range(n)]
    coords[-1][:, :, :2] += 500.0
    douts = [torch.randn(B, L * K * K, H, W, generator=g).to(dev()) for _ in range(n)]
    _cabi.check(lib.ezb_corr_lookup_bwd_multi(_cabi.ptr_array(douts), _cabi.ptr_array(coords), n, B, H, W, L, r, _cabi.ptr_array(lv), st))
    assert any(float(t.abs().max()) > 0 for t in lv)
    cnt = ctypes.c_int64(0)
    _cabi.check(lib.ezb_corr_grad_ranges_count(B, H, W, ctypes.byref(cnt)))
    rg = torch.empty(cnt.value, dtype=torch.int32, device=dev())
    _cabi.check(lib.ezb_corr_grad_ranges(_cabi.ptr_array(coords), n, B, H, W, L, r, _cabi.ptr(rg), st))
    _cabi.check(lib.ezb_corr_grad_rezero(_cabi.ptr_array(lv), _cabi.ptr(rg), B, H, W, L, st))
    torch.cuda.synchronize()
    for l, t in enumerate(lv):
        assert float(t.abs().max()) == 0.0, f"level {l}: the ranges miss part of the adjoint's support"


def test_pooled_gradient_pyramid_soak():
    """Twelve consecutive training-like steps through the pooled gradient pyramid with a different number of lookups, spread and
    offset each step (tight clusters, drifting windows, wide scatter, windows leaving the levels): every step's gradients
    must equal the dense adjoint's on a fresh zero-filled pyramid — any residue a step left behind would show up in the next."""
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    rng = np.random.default_rng(1234)
    B, C, H, W, L, r = 2, 16, 20, 28, 3, 4
    f1n = rng.standard_normal((B, C, H, W)).astype(np.float32)
    f2n = rng.standard_normal((B, C, H, W)).astype(np.float32)
    grid = O.coords_grid(B, H, W)
    E.release_gradient_pool()
    for step in range(12):
        n = int(rng.integers(1, 6))
        spread = float(rng.choice([0.2, 1.0, 4.0, 15.0]))
        shift = rng.uniform(-12, 12, (1, 2, 1, 1)).astype(np.float32) * (3.0 if step % 5 == 4 else 1.0)
        coords = [grid + shift + rng.normal(0, spread, (B, 2, H, W)).astype(np.float32) for _ in range(n)]
        douts = [rng.standard_normal((B, L * (2 * r + 1) ** 2, H, W)).astype(np.float32) for _ in range(n)]
        g1, g2 = _train_like_step(E, f1n, f2n, coords, douts, L, r)
        os.environ["EZB_BWD_DENSE"] = "1"
        os.environ["EZB_LOOKUP_BWD_SEPARATE"] = "1"
        try:
            d1, d2 = _train_like_step(E, f1n, f2n, coords, douts, L, r)
        finally:
            del os.environ["EZB_BWD_DENSE"], os.environ["EZB_LOOKUP_BWD_SEPARATE"]
        if float(d1.abs().max()) == 0.0:
            assert float(g1.abs().max()) == 0.0 and float(g2.abs().max()) == 0.0
            continue
        assert_close(g1, d1.cpu().numpy(), 3e-4, f"step {step} (n={n}, spread={spread}) dfmap1")
        assert_close(g2, d2.cpu().numpy(), 3e-4, f"step {step} (n={n}, spread={spread}) dfmap2")
    E.release_gradient_pool()


@pytest.mark.parametrize("store", ["fp16", "fp32"])
@pytest.mark.parametrize("cfg", [(1, 1, 21, 45), (2, 2, 33, 70), (5, 2, 30, 52), (7, 1, 40, 64), (8, 1, 36, 72)])
def test_lookup_radii_and_window_alignments_vs_oracle(cfg, store):
    """The gather fetches only the 16-byte chunks a window's 2r+2 columns span (their number travels with the window origin):
    every window start residue modulo the chunk width, radii from 1 to 8, the library's maximum (up to 6 chunks per row in fp32), windows hanging over every
    border, in both storage modes, against the oracle lookup (reference pairwise.py:43-69) — and the coords adjoint, which shares
    the gather."""
    import ezflow_b200 as E
    from oracle import similarity_oracle as O

    r, L, H, W = cfg
    rng = np.random.default_rng(100 + r)
    f1 = rng.standard_normal((1, 32, H, W)).astype(np.float32)
    f2 = rng.standard_normal((1, 32, H, W)).astype(np.float32)
    # integer offsets 0..15 along x make every chunk residue appear with every fraction; a margin beyond the borders
    coords = O.coords_grid(1, H, W)
    coords[:, 0] += (np.arange(W) % 16)[None, None, :] * 0.5 + rng.uniform(-9, 9, (1, H, W))
    coords[:, 1] += rng.uniform(-6, 6, (1, H, W))
    coords = coords.astype(np.float32)
    ref = O.corr_lookup(O.corr_pyramid(f1, f2, L), coords, r)
    E.set_pyramid_dtype(store)
    obj = E.MutliScalePairwise4DCorr(to_gpu(f1), to_gpu(f2), num_levels=L, corr_radius=r)
    out = obj(to_gpu(coords))
    assert tuple(out.shape) == ref.shape
    assert_close(out, ref, LOOK_TOL, f"lookup r={r} L={L} ({store})")
    if r <= 5:
        dout = rng.standard_normal(ref.shape).astype(np.float32)
        c = to_gpu(coords, True)
        (obj(c) * to_gpu(dout)).sum().backward()
        dref = O.corr_lookup_bwd_coords(O.corr_pyramid(f1, f2, L), coords, dout, r)
        assert_close(c.grad, dref, 5e-3, f"coords adjoint r={r} ({store})")
