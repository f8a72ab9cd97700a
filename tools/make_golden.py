#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference (neu-vi/ezflow,
imported from /root/reference through tests/refshim.py) on seeded inputs.

The reference cannot travel to the GPU box, so its outputs are committed as small
fixtures.  Re-run with:  python tools/make_golden.py [name-prefix ...]
Every fixture stores its inputs, so tests never depend on RNG reproducibility.
"""

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refshim  # noqa: E402

ez = refshim.import_reference()
assert ez is not None and refshim.reference_root() == "/root/reference", "needs /root/reference"

from ezflow.similarity import SIMILARITY_REGISTRY  # noqa: E402
from ezflow.similarity.correlation.sampler import iter_spatial_correlation_sample  # noqa: E402
from ezflow.utils import coords_grid  # noqa: E402
from ezflow.utils.resampling import convex_upsample_flow  # noqa: E402
from ezflow.utils.warp import warp  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)
torch.set_num_threads(4)


def npy(t):
    return t.detach().cpu().numpy()


ONLY = [a for a in sys.argv[1:] if not a.startswith("-")]  # optional fixture-name prefixes: regenerate only those


def save(name, **arrs):
    if ONLY and not any(name.startswith(p) for p in ONLY):
        return
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **{k: (npy(v) if torch.is_tensor(v) else np.asarray(v)) for k, v in arrs.items()})
    print(f"{name}: {os.path.getsize(path) / 1024:.0f} KiB")


def pairwise_case(name, shape, levels, radius, seed, jitter, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    b, c, h, w = shape
    f1 = (torch.randn(shape, generator=g) * scale).requires_grad_(True)
    f2 = (torch.randn(shape, generator=g) * scale).requires_grad_(True)
    coords = coords_grid(b, h, w) + (torch.rand(b, 2, h, w, generator=g) * 2 - 1) * jitter
    dout = torch.randn(b, levels * (2 * radius + 1) ** 2, h, w, generator=g)
    cls = SIMILARITY_REGISTRY.get("MutliScalePairwise4DCorr")
    obj = cls(f1, f2, num_levels=levels, corr_radius=radius)
    for p in obj.corr_pyramid:
        p.retain_grad()
    out = obj(coords)
    (out * dout).sum().backward()
    arrs = dict(
        f1=f1, f2=f2, coords=coords, dout=dout, out=out, df1=f1.grad, df2=f2.grad,
        levels=levels, radius=radius,
    )
    if b * h * w <= 200:
        arrs["corr"] = cls.corr(f1, f2)
    for i, p in enumerate(obj.corr_pyramid):
        arrs[f"pyr{i}"] = p
        arrs[f"dpyr{i}"] = p.grad
    save(name, **arrs)


def lookup_dcoords_case(name, shape, levels, radius, seed, jitter):
    """The lookup's gradient w.r.t. coords (RAFT detaches coords, but the reference's op produces it)."""
    g = torch.Generator().manual_seed(seed)
    b, c, h, w = shape
    f1 = torch.randn(shape, generator=g)
    f2 = torch.randn(shape, generator=g)
    coords = (coords_grid(b, h, w) + (torch.rand(b, 2, h, w, generator=g) * 2 - 1) * jitter).requires_grad_(True)
    dout = torch.randn(b, levels * (2 * radius + 1) ** 2, h, w, generator=g)
    obj = SIMILARITY_REGISTRY.get("MutliScalePairwise4DCorr")(f1, f2, num_levels=levels, corr_radius=radius)
    out = obj(coords)
    (out * dout).sum().backward()
    save(name, f1=f1, f2=f2, coords=coords, dout=dout, out=out, dcoords=coords.grad, levels=levels, radius=radius)


def sampler_case(name, shape, seed, **kw):
    g = torch.Generator().manual_seed(seed)
    x1 = torch.randn(shape, generator=g).requires_grad_(True)
    x2 = torch.randn(shape, generator=g).requires_grad_(True)
    out = iter_spatial_correlation_sample(x1, x2, **kw)
    dout = torch.randn(out.shape, generator=g)
    (out * dout).sum().backward()
    params = {k: np.asarray(v) for k, v in kw.items()}
    save(name, x1=x1, x2=x2, out=out, dout=dout, dx1=x1.grad, dx2=x2.grad, **{"p_" + k: v for k, v in params.items()})


def main():
    # ---- RAFT all-pairs correlation + pyramid + lookup (K1, K2, K3, K3b, K1b) ----
    pairwise_case("pairwise_l4_r4", (2, 32, 16, 21), 4, 4, seed=1, jitter=6.0)
    pairwise_case("pairwise_l3_r3", (1, 24, 10, 14), 3, 3, seed=2, jitter=3.0)
    pairwise_case("pairwise_scaled", (1, 64, 9, 12), 2, 2, seed=3, jitter=2.0, scale=30.0)
    lookup_dcoords_case("lookup_dcoords_l4_r4", (2, 16, 16, 21), 4, 4, seed=4, jitter=7.0)
    lookup_dcoords_case("lookup_dcoords_l2_r1", (1, 8, 9, 13), 2, 1, seed=5, jitter=12.0)

    # ---- local correlation (K4, K4b) ----
    sampler_case("sampler_p9", (2, 8, 13, 18), 11, patch_size=9)
    sampler_case("sampler_flownetc", (1, 16, 12, 20), 12, kernel_size=1, patch_size=21, padding=0, dilation_patch=2)
    sampler_case("sampler_stride2", (2, 6, 13, 17), 13, patch_size=5, stride=2)
    sampler_case("sampler_aniso", (2, 5, 11, 14), 14, patch_size=(3, 5), stride=(1, 2), padding=(1, 2), dilation_patch=(2, 1))
    sampler_case("sampler_dcv_s4", (1, 12, 24, 32), 15, patch_size=9, stride=4, dilation_patch=1)
    sampler_case("sampler_dil16", (1, 8, 12, 20), 16, patch_size=9, stride=1, dilation_patch=16)

    # ---- CorrelationLayer ----
    g = torch.Generator().manual_seed(21)
    a = torch.randn(2, 8, 12, 15, generator=g)
    b_ = torch.randn(2, 8, 12, 15, generator=g)
    save("corr_layer", x1=a, x2=b_, out=SIMILARITY_REGISTRY.get("CorrelationLayer")()(a, b_),
         out_p2_d3=SIMILARITY_REGISTRY.get("CorrelationLayer")(pad_size=3, max_displacement=3)(a, b_))

    # ---- Matryoshka (K6) ----
    g = torch.Generator().manual_seed(31)
    x1 = torch.randn(2, 16, 10, 14, generator=g)
    x2 = torch.randn(2, 16, 10, 14, generator=g)
    M = SIMILARITY_REGISTRY.get("MatryoshkaDilatedCostVolume")
    m0 = M()
    m1 = M(num_groups=4, max_displacement=2, stride=2, dilations=[1, 3], use_relu=True)
    save("matryoshka", x1=x1, x2=x2, out_default=m0(x1, x2), offsets_default=m0.get_relative_offsets(),
         out_g4=m1(x1, x2), offsets_g4=m1.get_relative_offsets())
    ML = SIMILARITY_REGISTRY.get("MatryoshkaDilatedCostVolumeList")
    xa = [torch.randn(1, 8, 32, 48, generator=g), torch.randn(1, 12, 8, 12, generator=g)]
    xb = [torch.randn(1, 8, 32, 48, generator=g), torch.randn(1, 12, 8, 12, generator=g)]
    l0 = ML()
    l1 = ML(normalize_feat_l2=True, use_relu=True)
    save("matryoshka_list", xa0=xa[0], xa1=xa[1], xb0=xb[0], xb1=xb[1], out_default=l0(xa, xb),
         out_norm_relu=l1(xa, xb), offsets_2d=l0.get_global_flow_offsets())

    # ---- warp (K5) ----
    g = torch.Generator().manual_seed(41)
    x = torch.randn(2, 6, 11, 15, generator=g).requires_grad_(True)
    flow = (torch.randn(2, 2, 11, 15, generator=g) * 3.0).requires_grad_(True)
    out = warp(x, flow)
    dout = torch.randn(out.shape, generator=g)
    (out * dout).sum().backward()
    save("warp", x=x, flow=flow, out=out, dout=dout, dx=x.grad, dflow=flow.grad)

    # ---- convex flow upsampling (K7) ----
    g = torch.Generator().manual_seed(51)
    for name, (n, c, h, w, st) in {"convex_up_s8": (1, 2, 4, 35, 8), "convex_up_s4": (2, 3, 5, 33, 4)}.items():
        flow = (torch.randn(n, c, h, w, generator=g) * 4.0).requires_grad_(True)
        logits = (torch.randn(n, 9 * st * st, h, w, generator=g) * 2.0).requires_grad_(True)
        out = convex_upsample_flow(flow, logits, st)
        dout = torch.randn(out.shape, generator=g)
        (out * dout).sum().backward()
        save(name, flow=flow, logits=logits, out=out, dout=dout, dflow=flow.grad, dlogits=logits.grad, out_stride=st)


def volume_cases():
    """VCN._corr_fn and the volume LearnableMatchingCost.forward feeds its matching net, forward + autograd."""
    from ezflow.models.vcn import VCN
    from ezflow.similarity.learnable_cost import LearnableMatchingCost

    g = torch.Generator().manual_seed(61)
    f1 = torch.randn(2, 5, 9, 12, generator=g).requires_grad_(True)
    f2 = torch.randn(2, 5, 9, 12, generator=g).requires_grad_(True)
    for name, (md, fac) in {"vcn_corr_md3": (3, 1), "vcn_corr_md4_f2": (4, 2)}.items():
        f1.grad = f2.grad = None
        out = VCN._corr_fn(None, f1, f2, md, fac)
        dout = torch.randn(out.shape, generator=g)
        (out * dout).sum().backward()
        save(name, f1=f1, f2=f2, out=out, dout=dout, df1=f1.grad, df2=f2.grad, max_disp=md, factorization=fac)

    class Capture(torch.nn.Module):  # stands in for the matching network: records the assembled volume
        def forward(self, cost):
            self.cost = cost
            return cost[:, :1]

    x = torch.randn(2, 4, 8, 11, generator=g)
    y = torch.randn(2, 4, 8, 11, generator=g)
    y[:, :, 2:4, 3:6] = 0.0  # warp holes
    x.requires_grad_(True)
    y.requires_grad_(True)
    for name, (mu, mv, holes) in {"dicl_volume_u3v3": (3, 3, True), "dicl_volume_u2v1_keep": (2, 1, False)}.items():
        x.grad = y.grad = None
        m = LearnableMatchingCost(max_u=mu, max_v=mv, remove_warp_hole=holes)
        m.matching_net = Capture()
        m(x, y)
        vol = m.matching_net.cost
        dout = torch.randn(vol.shape, generator=g)
        (vol * dout).sum().backward()
        save(name, x=x, y=y, out=vol, dout=dout, dx=x.grad, dy=y.grad, max_u=mu, max_v=mv, remove_warp_hole=int(holes))


if __name__ == "__main__":
    main()
    volume_cases()
