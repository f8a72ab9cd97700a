"""Helpers shared by the GPU parity tests (test infrastructure)."""

import math

import numpy as np
import torch
import torch.nn.functional as F


def dev():
    return torch.device("cuda:0")


def to_gpu(a, requires_grad=False):
    t = torch.from_numpy(np.ascontiguousarray(a)).to(dev())
    if requires_grad:
        t.requires_grad_(True)
    return t


def rel_fro(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def max_rel(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def assert_close(a, b, tol, what=""):
    """Both gates from SURVEY.md §8c: relative Frobenius error and max-abs / max-ref."""
    a = a.detach().cpu().numpy() if torch.is_tensor(a) else np.asarray(a)
    b = b.detach().cpu().numpy() if torch.is_tensor(b) else np.asarray(b)
    assert a.shape == b.shape, f"{what}: shape {a.shape} vs {b.shape}"
    assert np.isfinite(a).all(), f"{what}: non-finite values"
    rf, mr = rel_fro(a, b), max_rel(a, b)
    assert rf <= tol and mr <= tol, f"{what}: rel_fro={rf:.3e} max_rel={mr:.3e} > {tol:.1e}"
    return rf, mr


class TorchPairwiseRef:
    """fp32 PyTorch restatement of MutliScalePairwise4DCorr (reference ezflow/similarity/correlation/pairwise.py:26-88 with
    bilinear_sampler, ezflow/utils/resampling.py:56-85): the floating-point comparator for the CUDA path on the same GPU.
    `queries` (optional LongTensor of flat query indices, same for every batch entry) restricts the volume to those query
    rows, so the comparator also fits at 1080p (a full fp32 volume is 4.2 GB per pair)."""

    def __init__(self, fmap1, fmap2, num_levels, corr_radius, queries=None):
        b, c, h, w = fmap1.shape
        a = fmap1.view(b, c, h * w).transpose(1, 2)
        if queries is not None:
            a = a[:, queries]
        corr = torch.matmul(a, fmap2.view(b, c, h * w)) / math.sqrt(float(c))
        self.nq = a.shape[1]
        corr = corr.reshape(b * self.nq, 1, h, w)
        self.levels = [corr]
        for _ in range(num_levels - 1):
            self.levels.append(F.avg_pool2d(self.levels[-1], 2, stride=2))
        self.r = corr_radius
        self.queries = queries

    def __call__(self, coords):
        """coords (B,2,H,W) -> (B, L*(2r+1)^2, H, W), or (B, L*(2r+1)^2, nq) when built for a query subset."""
        b, _, h, w = coords.shape
        r = self.r
        c = coords.permute(0, 2, 3, 1).reshape(b, h * w, 2)
        if self.queries is not None:
            c = c[:, self.queries]
        c = c.reshape(b * self.nq, 1, 1, 2)
        d = torch.linspace(-r, r, 2 * r + 1, device=coords.device)
        delta = torch.stack(torch.meshgrid(d, d, indexing="ij"), dim=-1).view(1, 2 * r + 1, 2 * r + 1, 2)
        out = []
        for l, lvl in enumerate(self.levels):
            hl, wl = lvl.shape[-2:]
            pos = c / 2**l + delta  # the first window index offsets x (reference quirk, pairwise.py:53-61)
            gx = 2 * pos[..., 0] / (wl - 1) - 1
            gy = 2 * pos[..., 1] / (hl - 1) - 1
            s = F.grid_sample(lvl, torch.stack([gx, gy], dim=-1), align_corners=True)
            out.append(s.view(b, self.nq, -1))
        out = torch.cat(out, dim=-1).permute(0, 2, 1).contiguous().float()
        return out if self.queries is not None else out.view(b, -1, h, w)
