"""Hot-path helpers that sit next to the similarity ops in the reference.

* ``warp``        — reference ezflow/utils/warp.py:5-42 (PWC-Net feature warping, K5)
* ``coords_grid`` — reference ezflow/utils/common.py:9-31 (pixel grid feeding the RAFT lookup)
"""

import torch

from . import _cabi


class _WarpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, flow):
        ctx.save_for_backward(x, flow)
        return _warp_forward(x, flow)

    @staticmethod
    def backward(ctx, dout):
        x, flow = ctx.saved_tensors
        B, C, H, W = x.shape
        dout = dout.contiguous().float()
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dflow = torch.empty_like(flow) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(x.device):
            _cabi.check(
                _cabi.lib().ezb_warp_bwd(
                    _cabi.ptr(x), _cabi.ptr(flow), _cabi.ptr(dout), B, C, H, W, _cabi.ptr(dx), _cabi.ptr(dflow),
                    _cabi.stream_ptr(x.device),
                )
            )
        return dx, dflow


def _warp_forward(x, flow):
    B, C, H, W = x.shape
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _cabi.check(_cabi.lib().ezb_warp_fwd(_cabi.ptr(x), _cabi.ptr(flow), B, C, H, W, _cabi.ptr(out), _cabi.stream_ptr(x.device)))
    return out


def warp(x, flow):
    """Warps an image / feature map x (B,C,H,W) by the optical flow field flow (B,2,H,W)."""
    a = _cabi.require_cuda_f32(x, "x")
    f = _cabi.require_cuda_f32(flow, "flow")
    if a.dim() != 4 or f.dim() != 4 or f.shape[1] != 2 or f.shape[0] != a.shape[0] or f.shape[2:] != a.shape[2:]:
        raise ValueError(f"warp expects x (B,C,H,W) and flow (B,2,H,W); got {tuple(x.shape)} and {tuple(flow.shape)}")
    if torch.is_grad_enabled() and (a.requires_grad or f.requires_grad):
        return _WarpFn.apply(a, f)
    return _warp_forward(a, f)


def coords_grid(batch_size, h, w, device=None):
    """(batch_size, 2, h, w) grid, channel 0 = x index, channel 1 = y index."""
    ys, xs = torch.meshgrid(torch.arange(h, device=device), torch.arange(w, device=device), indexing="ij")
    coords = torch.stack([xs, ys], dim=0).float()
    return coords[None].repeat(batch_size, 1, 1, 1)
