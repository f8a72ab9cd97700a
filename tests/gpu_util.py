"""Helpers shared by the GPU parity tests (test infrastructure)."""

import numpy as np
import torch


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
