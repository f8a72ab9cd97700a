"""GPU parity AT THE BENCH CONFIGURATION: the exact launch bench.py times — 8 pairs of 256x135x240 feature maps in one
micro-batch, fp16 pyramid (23 GB; group-block offsets close to 2^31 elements), fractional coordinates drawn as in bench.py —
compared, for 2000 sampled queries of EVERY pair and all 324 output channels, with an fp32 computation on the same GPU:
f1[:, q]^T f2 / 16 (true fp32 matmul), floor 2x2 average pools, grid_sample taps (reference
ezflow/similarity/correlation/pairwise.py:26-69).  Tolerance: the north_star's 1e-3 relative (measured ~3e-4)."""

import pytest
import torch

from gpu_util import TorchPairwiseRef, dev

pytestmark = pytest.mark.gpu

H, W, C, L, R, P = 135, 240, 256, 4, 4, 8
TOL = 1e-3


def test_bench_config_b8_1080p_fractional_lookup_vs_fp32():
    import ezflow_b200 as E

    old = (torch.backends.cuda.matmul.allow_tf32, E.get_pyramid_dtype())
    torch.backends.cuda.matmul.allow_tf32 = False
    E.set_pyramid_dtype("fp16")
    try:
        g = torch.Generator(device="cpu").manual_seed(1234)
        f1 = torch.randn(P, C, H, W, generator=g).to(dev())
        f2 = torch.randn(P, C, H, W, generator=g).to(dev())
        grid = E.coords_grid(P, H, W)
        with torch.no_grad():
            fn = E.MutliScalePairwise4DCorr(f1, f2, num_levels=L, corr_radius=R)
            worst = 0.0
            for k in (0, 5, 11):  # first, a middle and the last of bench.py's 12 coordinate sets (noise sigma 2 + k pixels)
                coords = (grid + torch.randn(P, 2, H, W, generator=g) * (2 + k)).clamp_(-16, W + 16).to(dev())
                out = fn(coords)
                assert tuple(out.shape) == (P, L * (2 * R + 1) ** 2, H, W) and torch.isfinite(out).all()
                out = out.view(P, -1, H * W)
                for p in range(P):
                    # 2000 queries per pair: random ones plus the first / last rows and the corners (window partly outside)
                    q = torch.randperm(H * W, generator=g)[:1980]
                    q = torch.cat([q, torch.arange(0, 5), torch.arange(W - 5, W), torch.arange(H * W - 10, H * W)]).to(dev())
                    ref = TorchPairwiseRef(f1[p : p + 1], f2[p : p + 1], L, R, queries=q)(coords[p : p + 1])[0]  # (324, 2000)
                    got = out[p][:, q]
                    scale = ref.abs().max().item()
                    max_rel = (got - ref).abs().max().item() / scale
                    rel_fro = ((got - ref).norm() / ref.norm()).item()
                    worst = max(worst, max_rel, rel_fro)
                    assert max_rel <= TOL and rel_fro <= TOL, f"pair {p}, coords set {k}: max_rel {max_rel:.3e} rel_fro {rel_fro:.3e}"
                    # per level, so a small level cannot hide behind level 0's magnitude
                    for l in range(L):
                        a, b = got[l * 81 : (l + 1) * 81], ref[l * 81 : (l + 1) * 81]
                        lr = ((a - b).norm() / b.norm()).item()
                        assert lr <= TOL, f"pair {p}, coords set {k}, level {l}: rel_fro {lr:.3e}"
                del out
        print(f"bench-config parity: 8 pairs x 3 coordinate sets x 2000 queries x 324 channels, worst relative error {worst:.2e}")
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old[0]
        E.set_pyramid_dtype(old[1])
        torch.cuda.empty_cache()
