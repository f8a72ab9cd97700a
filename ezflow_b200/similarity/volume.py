"""SURVEY.md §8f-4 on B200: VCN's per-channel correlation volume, DICL's displacement-stacked matching volume, and the
memory-free ("on-demand") variant of RAFT's correlation lookup.

* ``vcn_corr(features1, features2, max_disp, factorization=1)`` — ``VCN._corr_fn`` (reference ezflow/models/vcn.py:162-206);
  ``install()`` rebinds the method.
* ``dicl_cost_volume(x, y, max_u, max_v, remove_warp_hole)`` — the volume assembly of ``LearnableMatchingCost.forward``
  (reference ezflow/similarity/learnable_cost.py:181-226), returned in the layout the matching network consumes
  (B*U*V, 2C, H, W); ``install()`` swaps a subclass of the reference's ``LearnableMatchingCost`` whose ``forward`` uses it.
* ``OnDemandPairwise4DCorr`` — same constructor / call surface as ``MutliScalePairwise4DCorr`` (pairwise.py:9-88) but no
  all-pairs volume: every lookup forms the (2r+2)^2 dot products per query and level from the feature maps (fp32, exact).
  Memory O(N*C) instead of O(N^2); ~30x slower per lookup than the pyramid path at 1080p, for inputs whose pyramid would not fit.
  Inference only (no autograd).
"""

import ctypes

import torch

from .. import _cabi
from ..config import configurable


class _VcnCorrFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, f1, f2, mdx, mdy, slope):
        B, C, H, W = f1.shape
        out = torch.empty((B, C, 2 * mdx + 1, 2 * mdy + 1, H, W), dtype=torch.float32, device=f1.device)
        if out.numel():
            with torch.cuda.device(f1.device):
                _cabi.check(_cabi.lib().ezb_vcn_corr_fwd(_cabi.ptr(f1), _cabi.ptr(f2), B, C, H, W, mdx, mdy, float(slope), _cabi.ptr(out),
                                                         _cabi.stream_ptr(f1.device)))
        ctx.cfg = (mdx, mdy, slope)
        ctx.save_for_backward(f1, f2)
        return out

    @staticmethod
    def backward(ctx, dout):
        f1, f2 = ctx.saved_tensors
        mdx, mdy, slope = ctx.cfg
        B, C, H, W = f1.shape
        d1 = torch.empty_like(f1) if ctx.needs_input_grad[0] else None
        d2 = torch.empty_like(f2) if ctx.needs_input_grad[1] else None
        if f1.numel() and (d1 is not None or d2 is not None):
            g = dout.contiguous().float()
            with torch.cuda.device(f1.device):
                _cabi.check(_cabi.lib().ezb_vcn_corr_bwd(_cabi.ptr(f1), _cabi.ptr(f2), _cabi.ptr(g), B, C, H, W, mdx, mdy, float(slope),
                                                         _cabi.ptr(d1), _cabi.ptr(d2), _cabi.stream_ptr(f1.device)))
        return d1, d2, None, None, None


def vcn_corr(features1, features2, max_disp, factorization=1, negative_slope=0.1):
    """(B, C, 2*max_disp+1, 2*(max_disp//factorization)+1, H, W) per-channel shifted products with LeakyReLU(0.1)."""
    a = _cabi.require_cuda_f32(features1, "features1")
    b = _cabi.require_cuda_f32(features2, "features2")
    if a.dim() != 4 or a.shape != b.shape:
        raise ValueError(f"features1/features2 must both be (B,C,H,W); got {tuple(features1.shape)} and {tuple(features2.shape)}")
    return _VcnCorrFn.apply(a, b, int(max_disp), int(max_disp // factorization), float(negative_slope))


class _DiclVolumeFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, y, max_u, max_v, remove_warp_hole):
        B, C, H, W = x.shape
        U, V = 2 * max_u + 1, 2 * max_v + 1
        out = torch.empty((B * U * V, 2 * C, H, W), dtype=torch.float32, device=x.device)
        need_mask = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]  # the keep-mask is all the backward needs
        mask = torch.empty((B * U * V, H, W), dtype=torch.uint8, device=x.device) if need_mask else None
        if out.numel():
            with torch.cuda.device(x.device):
                _cabi.check(_cabi.lib().ezb_dicl_volume_fwd(_cabi.ptr(x), _cabi.ptr(y), B, C, H, W, max_u, max_v, 1 if remove_warp_hole else 0,
                                                            _cabi.ptr(out), _cabi.ptr(mask), _cabi.stream_ptr(x.device)))
        ctx.cfg = (B, C, H, W, max_u, max_v)
        ctx.mask = mask
        return out

    @staticmethod
    def backward(ctx, dout):
        B, C, H, W, max_u, max_v = ctx.cfg
        dx = torch.empty((B, C, H, W), dtype=torch.float32, device=dout.device) if ctx.needs_input_grad[0] else None
        dy = torch.empty((B, C, H, W), dtype=torch.float32, device=dout.device) if ctx.needs_input_grad[1] else None
        if dout.numel() and (dx is not None or dy is not None):
            g = dout.contiguous().float()
            with torch.cuda.device(dout.device):
                _cabi.check(_cabi.lib().ezb_dicl_volume_bwd(_cabi.ptr(g), _cabi.ptr(ctx.mask), B, C, H, W, max_u, max_v, _cabi.ptr(dx), _cabi.ptr(dy),
                                                            _cabi.stream_ptr(dout.device)))
        return dx, dy, None, None, None


def dicl_cost_volume(x, y, max_u, max_v, remove_warp_hole=True):
    """(B*U*V, 2C, H, W): for every displacement (i, j) the stacked [x ; shifted y] image the matching net scores."""
    a = _cabi.require_cuda_f32(x, "x")
    b = _cabi.require_cuda_f32(y, "y")
    if a.dim() != 4 or a.shape != b.shape:
        raise ValueError(f"x/y must both be (B,C,H,W); got {tuple(x.shape)} and {tuple(y.shape)}")
    return _DiclVolumeFn.apply(a, b, int(max_u), int(max_v), bool(remove_warp_hole))


def learnable_matching_cost_forward(self, x, y):
    """``LearnableMatchingCost.forward`` (reference learnable_cost.py:157-231) with the volume assembly as one kernel; the
    matching network and the output views are the reference's."""
    if getattr(self, "cuda_cost_compute", False):  # the reference's own alternative: a plain correlation (learnable_cost.py:174-183)
        from .sampler import IterSpatialCorrelationSampler

        return IterSpatialCorrelationSampler(kernel_size=1, patch_size=(7, 7), stride=1, padding=0, dilation_patch=1)(x, y)
    size_u, size_v = 2 * self.max_u + 1, 2 * self.max_v + 1
    cost = dicl_cost_volume(x, y, self.max_u, self.max_v, self.remove_warp_hole)
    cost = self.matching_net(cost)
    cost = cost.view(x.size()[0], size_u, size_v, 1, x.size()[2], x.size()[3])
    return cost.permute([0, 3, 1, 2, 4, 5]).contiguous()


class OnDemandPairwise4DCorr:
    """
    RAFT correlation lookups without the all-pairs volume (sm_100a, fp32).

    Parameters
    ----------
    fmap1, fmap2 : torch.Tensor
        (B, C, H, W) CUDA feature maps
    num_levels : int
        Number of levels in the (virtual) correlation pyramid, at most 4
    corr_radius : int
        Radius of the lookup window
    """

    @configurable
    def __init__(self, fmap1, fmap2, num_levels=4, corr_radius=4):
        self.num_levels = num_levels
        self.corr_radius = corr_radius
        f1 = _cabi.require_cuda_f32(fmap1, "fmap1").detach()
        f2 = _cabi.require_cuda_f32(fmap2, "fmap2").detach()
        if f1.dim() != 4 or f1.shape != f2.shape:
            raise ValueError(f"fmap1/fmap2 must both be (B,C,H,W); got {tuple(fmap1.shape)} and {tuple(fmap2.shape)}")
        self._shape = tuple(f1.shape)
        self._ws = None
        B, C, H, W = self._shape
        if B == 0:
            return
        lib = _cabi.lib()
        n = ctypes.c_size_t(0)
        _cabi.check(lib.ezb_corr_ondemand_workspace_bytes(B, C, H, W, int(num_levels), ctypes.byref(n)))
        self._ws = torch.empty(n.value, dtype=torch.uint8, device=f1.device)
        with torch.cuda.device(f1.device):
            _cabi.check(lib.ezb_corr_ondemand_prepare(_cabi.ptr(f1), _cabi.ptr(f2), B, C, H, W, int(num_levels), _cabi.ptr(self._ws), n.value,
                                                      _cabi.stream_ptr(f1.device)))

    @classmethod
    def from_config(cls, cfg):
        return {"num_levels": cfg.NUM_LEVELS, "corr_radius": cfg.CORR_RADIUS}

    def __call__(self, coords):
        c = _cabi.require_cuda_f32(coords, "coords").detach()
        B, C, H, W = self._shape
        k = 2 * int(self.corr_radius) + 1
        if c.dim() != 4 or tuple(c.shape) != (B, 2, H, W):
            raise ValueError(f"coords must be ({B},2,{H},{W}); got {tuple(coords.shape)}")
        out = torch.empty((B, int(self.num_levels) * k * k, H, W), dtype=torch.float32, device=c.device)
        if B == 0:
            return out
        with torch.cuda.device(c.device):
            _cabi.check(_cabi.lib().ezb_corr_ondemand_lookup(_cabi.ptr(self._ws), _cabi.ptr(c), B, C, H, W, int(self.num_levels),
                                                             int(self.corr_radius), _cabi.ptr(out), _cabi.stream_ptr(c.device)))
        return out
