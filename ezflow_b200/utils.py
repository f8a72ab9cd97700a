"""Hot-path helpers that sit next to the similarity ops in the reference.

* ``warp``        — reference ezflow/utils/warp.py:5-42 (PWC-Net feature warping, K5)
* ``coords_grid`` — reference ezflow/utils/common.py:9-31 (pixel grid feeding the RAFT lookup)
* ``convex_upsample_flow`` — reference ezflow/utils/resampling.py:112-141 (consumer of every RAFT / DCVNet
  iteration's coarse flow, models/raft.py:127, decoder/dilated_flow_stack_filter.py:219; K7)
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
        return _warp_backward(x, flow, dout.contiguous().float(), ctx.needs_input_grad[0], ctx.needs_input_grad[1])


def _warp_backward(x, flow, dout, need_x=True, need_flow=True):
    B, C, H, W = x.shape
    dx = torch.empty_like(x) if need_x else None
    dflow = torch.empty_like(flow) if need_flow else None
    if dx is None and dflow is None:
        return None, None
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
    if B == 0:
        return out
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


class _ConvexUpsampleFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, flow, mask_logits, out_stride):
        ctx.save_for_backward(flow, mask_logits)
        ctx.out_stride = out_stride
        return _convex_upsample_forward(flow, mask_logits, out_stride)

    @staticmethod
    def backward(ctx, dout):
        flow, logits = ctx.saved_tensors
        N, C, H, W = flow.shape
        dout = dout.contiguous().float()
        dflow = torch.empty_like(flow) if ctx.needs_input_grad[0] else None
        dlogits = torch.empty_like(logits) if ctx.needs_input_grad[1] else None
        lib = _cabi.lib()
        with torch.cuda.device(flow.device):
            ws, nbytes = None, 0
            if dflow is not None:
                need = _cabi.c_sz()
                _cabi.check(lib.ezb_convex_upsample_bwd_workspace_bytes(N, C, H, W, _cabi.ctypes.byref(need)))
                nbytes = need.value
                ws = torch.empty(nbytes, dtype=torch.uint8, device=flow.device)
            _cabi.check(
                lib.ezb_convex_upsample_bwd(
                    _cabi.ptr(flow), _cabi.ptr(logits), _cabi.ptr(dout), N, C, H, W, ctx.out_stride, _cabi.ptr(dflow),
                    _cabi.ptr(dlogits), _cabi.ptr(ws), nbytes, _cabi.stream_ptr(flow.device),
                )
            )
        return dflow, dlogits, None


def _convex_upsample_forward(flow, mask_logits, out_stride):
    N, C, H, W = flow.shape
    out = torch.empty((N, C, out_stride * H, out_stride * W), dtype=torch.float32, device=flow.device)
    if N == 0:
        return out
    with torch.cuda.device(flow.device):
        _cabi.check(
            _cabi.lib().ezb_convex_upsample_fwd(
                _cabi.ptr(flow), _cabi.ptr(mask_logits), N, C, H, W, out_stride, _cabi.ptr(out), _cabi.stream_ptr(flow.device)
            )
        )
    return out


def convex_upsample_flow(flow, mask_logits, out_stride):
    """Upsamples flow (N,C,H,W) to (N,C,s*H,s*W) as a softmax-weighted combination of each pixel's 3x3 coarse
    neighbourhood; mask_logits holds N x (9*s*s) x H x W logits (viewed (N,1,9,s,s,H,W) like the reference)."""
    f = _cabi.require_cuda_f32(flow, "flow")
    m = _cabi.require_cuda_f32(mask_logits, "mask_logits")
    s = int(out_stride)
    if f.dim() != 4 or m.dim() != 4 or s <= 0 or m.shape[0] != f.shape[0] or m.shape[1] != 9 * s * s or m.shape[2:] != f.shape[2:]:
        raise RuntimeError(
            f"convex_upsample_flow expects flow (N,C,H,W) and mask_logits (N,9*{s}*{s},H,W); got {tuple(flow.shape)} and "
            f"{tuple(mask_logits.shape)}"
        )  # the reference's .view raises RuntimeError for a mismatched mask
    if torch.is_grad_enabled() and (f.requires_grad or m.requires_grad):
        return _ConvexUpsampleFn.apply(f, m, s)
    return _convex_upsample_forward(f, m, s)
