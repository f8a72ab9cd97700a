#!/usr/bin/env python
"""End-to-end RAFT parity fixture (BASELINE config 1: raft.yaml, random init, 368x496, 12 iterations).

Runs the UNMODIFIED reference RAFT (imported from /root/reference through tests/refshim.py) on CPU and writes

  tests/golden/raft_e2e.npz        images (uint8) + the reference's final flow (committed, small)
  oracle/_ref/raft_cfg1_parts.pt   TorchScript traces of the reference's NON-similarity parts with the seeded
                                   random-init weights (encoders; update block + convex upsampling).  Git-ignored,
                                   travels to the GPU box like a built .so (it is an artefact derived from the
                                   reference, produced by this committed recipe; no reference source is copied).

tests/test_gpu_raft_e2e.py then runs the 12-iteration RAFT loop on the GPU with the traced parts around (a) a plain
PyTorch fp32 restatement of the similarity path and (b) this repo's CUDA path, and checks mean EPE <= 0.01 px
(north_star).  Re-run with:  python tools/make_raft_e2e.py
"""

import os
import sys

import numpy as np
import torch
import torch.nn as nn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import refshim  # noqa: E402

H, W, ITERS = 368, 496, 12


def smooth_pair(seed=1234):
    """Low-pass-filtered noise and a copy shifted by a few pixels, as 8-bit images (flows are non-trivial)."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(1, 3, H // 4 + 8, W // 4 + 8, generator=g)
    big = torch.nn.functional.interpolate(base, scale_factor=4, mode="bicubic", align_corners=False)
    img1 = big[:, :, 16 : 16 + H, 16 : 16 + W]
    img2 = big[:, :, 13 : 13 + H, 19 : 19 + W]  # shifted by (+3 rows, -3 cols)
    q = lambda t: (t * 48 + 128).round().clamp(0, 255).to(torch.uint8)
    return q(img1), q(img2)


def main():
    ez = refshim.import_reference()
    assert ez is not None and refshim.reference_root() == "/root/reference", "needs /root/reference"
    from ezflow.models import build_model
    from ezflow.utils import convex_upsample_flow

    torch.set_num_threads(8)
    torch.manual_seed(0)
    model = build_model("RAFT", cfg_path=refshim.reference_config_path("raft.yaml"), custom_cfg=True).eval()
    assert model.cfg.UPDATE_ITERS == ITERS and not model.cfg.MIXED_PRECISION

    u1, u2 = smooth_pair()
    img1, img2 = u1.float(), u2.float()
    with torch.no_grad():
        out = model(img1, img2)
    flow = out["flow_upsampled"]
    print("reference flow: mean |flow| = %.3f px, max %.3f" % (flow.norm(dim=1).mean().item(), flow.abs().max().item()))

    class Front(nn.Module):  # reference models/raft.py:88-109 minus the similarity construction
        def __init__(self, m):
            super().__init__()
            self.fnet, self.cnet, self.hd, self.cd = m.fnet, m.cnet, m.cfg.HIDDEN_DIM, m.cfg.CONTEXT_DIM

        def forward(self, a, b):
            fmap1, fmap2 = self.fnet([a, b])
            cnet = self.cnet(a)
            net, inp = torch.split(cnet, [self.hd, self.cd], dim=1)
            return fmap1.float(), fmap2.float(), torch.tanh(net), torch.relu(inp)

    class Step(nn.Module):  # reference models/raft.py:118-129 (everything after corr_fn(coords1))
        def __init__(self, m):
            super().__init__()
            self.update_block = m.update_block

        def forward(self, net, inp, corr, coords0, coords1):
            flow = coords1 - coords0
            net, up_mask, delta_flow = self.update_block(net, inp, corr, flow)
            coords1 = coords1 + delta_flow
            flow_up = convex_upsample_flow(8 * (coords1 - coords0), up_mask, out_stride=8)
            return net, coords1, flow_up

    front, step = Front(model).eval(), Step(model).eval()
    with torch.no_grad():
        f1, f2, net, inp = front(img1, img2)
        h8, w8 = H // 8, W // 8
        from ezflow.utils import coords_grid

        c0 = coords_grid(1, h8, w8)
        corr = torch.zeros(1, model.corr_levels * (2 * model.corr_radius + 1) ** 2, h8, w8)
        t_front = torch.jit.trace(front, (img1, img2), check_trace=False)
        t_step = torch.jit.trace(step, (net, inp, corr, c0, c0.clone()), check_trace=False)
        # self-check: the traced parts + the reference's own similarity class reproduce the model's output
        from ezflow.similarity import SIMILARITY_REGISTRY

        cls = SIMILARITY_REGISTRY.get("MutliScalePairwise4DCorr")
        g1, g2, n, i = t_front(img1, img2)
        fn = cls(g1, g2, num_levels=model.corr_levels, corr_radius=model.corr_radius)
        c1 = c0.clone()
        for _ in range(ITERS):
            n, c1, up = t_step(n, i, fn(c1), c0, c1)
        epe = (up - flow).norm(dim=1).mean().item()
        print("traced parts + reference similarity vs model: mean EPE %.2e px" % epe)
        assert epe < 1e-4

    os.makedirs(os.path.join(ROOT, "oracle", "_ref"), exist_ok=True)
    parts = os.path.join(ROOT, "oracle", "_ref", "raft_cfg1_parts.pt")
    holder = nn.Module()
    holder.front, holder.step = t_front, t_step
    torch.jit.save(torch.jit.script(holder), parts)
    print("wrote", parts, "%.1f MB" % (os.path.getsize(parts) / 1e6))
    gold = os.path.join(ROOT, "tests", "golden", "raft_e2e.npz")
    np.savez_compressed(gold, img1=u1.numpy(), img2=u2.numpy(), flow=flow.numpy().astype(np.float32),
                        levels=model.corr_levels, radius=model.corr_radius, iters=ITERS)
    print("wrote", gold, "%.0f KiB" % (os.path.getsize(gold) / 1024))


if __name__ == "__main__":
    main()
