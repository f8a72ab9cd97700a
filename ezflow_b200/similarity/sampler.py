"""Local displacement-window correlation on B200 (K4 / K4b).

Drop-in for ``iter_spatial_correlation_sample`` / ``IterSpatialCorrelationSampler`` (reference
ezflow/similarity/correlation/sampler.py:29-209): same parameters, same ``NotImplementedError`` for
kernel_size != 1, dilation != 1 and even patch sizes, output (B, P_h, P_w, h', w') contiguous fp32.
"""

import ctypes
import os
from typing import Tuple, Union

import torch
import torch.nn as nn

from .. import _cabi
from ..config import configurable
from ..registry import SIMILARITY_REGISTRY

IntOrPair = Union[int, Tuple[int, int]]


def _pair(v):
    return (v, v) if isinstance(v, int) else (int(v[0]), int(v[1]))


def _out_hw(H, W, stride, padding):
    return -(-(H + 2 * padding[0]) // stride[0]), -(-(W + 2 * padding[1]) // stride[1])


_LOCAL_CORR_MODE = os.environ.get("EZB_LOCAL_CORR", "tc").lower()


def set_local_corr_mode(name):
    """'tc' (default): tcgen05 tensor cores, fp16 operands / fp32 accumulation, <= 1e-3 relative error;
    'exact': fp32 CUDA-core kernel (bit-for-bit fp32 FMA sums)."""
    global _LOCAL_CORR_MODE
    name = str(name).lower()
    if name not in ("tc", "exact"):
        raise ValueError("local correlation mode must be 'tc' or 'exact'")
    _LOCAL_CORR_MODE = name


def get_local_corr_mode():
    return _LOCAL_CORR_MODE


_WS_BYTES = {}  # (BG, C, H, W, padh, padw) -> bytes: the size query is a pure function of the shape


def local_corr_workspace(BG, C, H, W, device, padding=(0, 0)):
    """Scratch for the tensor-core path (fp16 K-major operand copies + scales), or (None, 0) in 'exact' mode."""
    if _LOCAL_CORR_MODE != "tc":
        return None, 0
    key = (BG, C, H, W, int(padding[0]), int(padding[1]))
    nbytes = _WS_BYTES.get(key)
    if nbytes is None:
        n = ctypes.c_size_t(0)
        _cabi.check(_cabi.lib().ezb_local_corr_workspace_bytes(*key, ctypes.byref(n)))
        nbytes = _WS_BYTES[key] = n.value
    return torch.empty(nbytes, dtype=torch.uint8, device=device), nbytes


TC_MAX_CHANNELS = 256  # per group; above it (or in 'exact' mode) the fp32 CUDA-core kernel runs


def warp_fusable(C):
    """True when ``corr(in1, warp(in2, flow))`` can run as one fused pass (tensor-core path only)."""
    return _LOCAL_CORR_MODE == "tc" and C <= TC_MAX_CHANNELS


def local_corr_forward(in1, in2, patch, stride, padding, dpatch, scale=1.0, slope=1.0, out=None, out_batch_stride=0, flow2=None):
    """out = act(scale * corr(in1, in2)); with ``flow2`` (B,2,H,W) the targets are warp(in2, flow2), evaluated on the fly."""
    B, C, H, W = in1.shape
    oh, ow = _out_hw(H, W, stride, padding)
    if out is None:
        out = torch.empty((B, patch[0], patch[1], oh, ow), dtype=torch.float32, device=in1.device)
    if B == 0:
        return out
    lib = _cabi.lib()
    with torch.cuda.device(in1.device):
        ws, nws = local_corr_workspace(B, C, H, W, in1.device, padding)
        geom = (B, C, H, W, patch[0], patch[1], stride[0], stride[1], padding[0], padding[1], dpatch[0], dpatch[1])
        tail = (float(scale), float(slope), _cabi.ptr(out), int(out_batch_stride), _cabi.ptr(ws), nws, _cabi.stream_ptr(in1.device))
        if flow2 is not None:
            _cabi.check(lib.ezb_warp_local_corr_fwd(_cabi.ptr(in1), _cabi.ptr(in2), _cabi.ptr(flow2), *geom, *tail))
        else:
            _cabi.check(lib.ezb_local_corr_fwd(_cabi.ptr(in1), _cabi.ptr(in2), *geom, *tail))
    return out


def bwd_tc_covers(B, C, H, W, patch, stride, padding, dpatch):
    """True when the tensor-core adjoint (csrc/local_corr_bwd_tc.cu) both handles the problem — the regular case; the library
    has the last word — and pays off: it has ~0.1 ms of fixed cost (three scale reductions, two operand conversions) plus ~11.5 ns
    per million gradient elements, against ~0.31 ms per 1e9 FMAs of the fp32 kernels (tools/micro/k4b_levels.py on B200:
    FlowNetC 0.88 -> 0.17 ms, DCVNet's stride-1 dilations 0.16-0.24 -> 0.10-0.14 ms; PWC-Net's levels stay on the fp32 kernels,
    which for that shape are the register-tiled ones: finest level 0.25 -> 0.155 ms)."""
    d = dpatch[0]
    regular = (
        _LOCAL_CORR_MODE == "tc" and C <= TC_MAX_CHANNELS and patch[0] == patch[1] and patch[0] % 2 == 1 and patch[0] <= 25
        and stride == (1, 1) and padding == (0, 0) and dpatch[0] == dpatch[1] and 1 <= d <= 8
        and (d == 1 or (-(-H // d) >= 8 and -(-W // d) >= 16))
    )
    elems = B * patch[0] * patch[1] * H * W  # gradient elements
    # fp32 side: 0.31 ms per 1e9 FMAs for the tiled kernels, 0.10 where the register-tiled adjoints of csrc/local_corr_rt.cu
    # apply (window <= 9, no dilation, W % 4 == 0: PWC-Net, DCVNet's dilation-1 volume)
    fp32_cost = 0.10 if (patch[0] <= 9 and d == 1 and W % 4 == 0) else 0.31
    return regular and elems * (fp32_cost * C - 11.5) > 65e6


def local_corr_backward(in1, in2, dout, patch, stride, padding, dpatch, scale=1.0, need1=True, need2=True, dout_batch_stride=0):
    """Both adjoints of the local correlation.  'tc' mode: banded tcgen05 GEMMs (fp16 operands, <= 1e-3 relative) where the
    problem is regular (stride 1, no padding, square window, one dilation); otherwise, and in 'exact' mode, fp32 CUDA-core kernels."""
    B, C, H, W = in1.shape
    d1 = torch.empty_like(in1) if need1 else None
    d2 = torch.empty_like(in2) if need2 else None
    if B == 0 or (d1 is None and d2 is None):
        return d1, d2
    lib = _cabi.lib()
    geom = (B, C, H, W, patch[0], patch[1], stride[0], stride[1], padding[0], padding[1], dpatch[0], dpatch[1])
    with torch.cuda.device(in1.device):
        if bwd_tc_covers(B, C, H, W, patch, stride, padding, dpatch):
            n = ctypes.c_size_t(0)
            _cabi.check(lib.ezb_local_corr_bwd_workspace_bytes(B, C, H, W, ctypes.byref(n)))
            ws = torch.empty(n.value, dtype=torch.uint8, device=in1.device)
            _cabi.check(lib.ezb_local_corr_bwd_ex(_cabi.ptr(in1), _cabi.ptr(in2), _cabi.ptr(dout), *geom, float(scale), int(dout_batch_stride),
                                                  _cabi.ptr(d1), _cabi.ptr(d2), _cabi.ptr(ws), n.value, _cabi.stream_ptr(in1.device)))
        else:
            _cabi.check(lib.ezb_local_corr_bwd(_cabi.ptr(in1), _cabi.ptr(in2), _cabi.ptr(dout), *geom, float(scale), int(dout_batch_stride),
                                               _cabi.ptr(d1), _cabi.ptr(d2), _cabi.stream_ptr(in1.device)))
    return d1, d2


class _LocalCorrFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, in1, in2, patch, stride, padding, dpatch, scale):
        ctx.cfg = (patch, stride, padding, dpatch, scale)
        ctx.save_for_backward(in1, in2)
        return local_corr_forward(in1, in2, patch, stride, padding, dpatch, scale)

    @staticmethod
    def backward(ctx, dout):
        in1, in2 = ctx.saved_tensors
        patch, stride, padding, dpatch, scale = ctx.cfg
        d1, d2 = local_corr_backward(
            in1, in2, dout.contiguous().float(), patch, stride, padding, dpatch, scale, ctx.needs_input_grad[0], ctx.needs_input_grad[1]
        )
        return d1, d2, None, None, None, None, None


class _FusedLocalCorrFn(torch.autograd.Function):
    """act(scale * corr(in1, warp(in2, flow))) with the warp (optional) and the LeakyReLU folded into the forward kernel.
    The backward chains the existing adjoints: activation mask -> K4b -> warp backward (the warped tensor is
    re-materialised there; the forward never writes it)."""

    @staticmethod
    def forward(ctx, in1, in2, flow, patch, stride, padding, dpatch, scale, slope):
        ctx.cfg = (patch, stride, padding, dpatch, scale, slope)
        out = local_corr_forward(in1, in2, patch, stride, padding, dpatch, scale, slope, flow2=flow)
        ctx.has_flow = flow is not None
        saved = [in1, in2] + ([flow] if ctx.has_flow else []) + ([out] if slope != 1.0 else [])
        ctx.save_for_backward(*saved)
        return out

    @staticmethod
    def backward(ctx, dout):
        from ..utils import _warp_backward, _warp_forward

        patch, stride, padding, dpatch, scale, slope = ctx.cfg
        saved = list(ctx.saved_tensors)
        in1, in2 = saved[0], saved[1]
        flow = saved[2] if ctx.has_flow else None
        dout = dout.contiguous().float()
        if slope != 1.0:  # slope > 0, so the sign of the activated output is the sign of the pre-activation
            out = saved[-1]
            dout = torch.where(out > 0, dout, dout * slope)
        need1, need2, needf = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.has_flow and ctx.needs_input_grad[2]
        tgt = _warp_forward(in2, flow) if ctx.has_flow else in2
        d1, dt = local_corr_backward(in1, tgt, dout, patch, stride, padding, dpatch, scale, need1, need2 or needf)
        d2, dflow = dt, None
        if ctx.has_flow and dt is not None:
            d2, dflow = _warp_backward(in2, flow, dt, need2, needf)
        return d1, d2, dflow, None, None, None, None, None, None


class WarpedFeatures:
    """Deferred ``warp(x, flow)`` (reference ezflow/utils/warp.py:5-42).  The PWC-Net decoder warps features2 only to hand
    them to the correlation sampler (reference ezflow/decoder/pyramid.py:138-141); passing this handle as ``input2`` lets the
    sampler evaluate the warp inside its operand conversion.  ``materialize()`` gives the ordinary warped tensor."""

    def __init__(self, x, flow):
        self.x = _cabi.require_cuda_f32(x, "x")
        self.flow = _cabi.require_cuda_f32(flow, "flow")
        f, a = self.flow, self.x
        if a.dim() != 4 or f.dim() != 4 or f.shape[1] != 2 or f.shape[0] != a.shape[0] or f.shape[2:] != a.shape[2:]:
            raise ValueError(f"warp expects x (B,C,H,W) and flow (B,2,H,W); got {tuple(x.shape)} and {tuple(flow.shape)}")

    shape = property(lambda self: self.x.shape)
    device = property(lambda self: self.x.device)
    dtype = property(lambda self: self.x.dtype)

    def dim(self):
        return self.x.dim()

    def materialize(self):
        from ..utils import warp

        return warp(self.x, self.flow)


def lazy_warp(x, flow):
    """Same signature as ``warp``; returns a ``WarpedFeatures`` handle for the sampler to consume."""
    return WarpedFeatures(x, flow)


def _validate(kernel_size, patch_size, dilation):
    # same checks, same messages as sampler.py:89-94
    if kernel_size[0] != 1 or kernel_size[1] != 1:
        raise NotImplementedError("Only kernel_size=1 is supported.")
    if dilation[0] != 1 or dilation[1] != 1:
        raise NotImplementedError("Only dilation=1 is supported.")
    if (patch_size[0] % 2) == 0 or (patch_size[1] % 2) == 0:
        raise NotImplementedError("Only odd patch sizes are supperted.")


def iter_spatial_correlation_sample(
    input1: torch.Tensor,
    input2: torch.Tensor,
    kernel_size: IntOrPair = 1,
    patch_size: IntOrPair = 1,
    stride: IntOrPair = 1,
    padding: IntOrPair = 0,
    dilation: IntOrPair = 1,
    dilation_patch: IntOrPair = 1,
    _scale: float = 1.0,
    _negative_slope: float = None,
) -> torch.Tensor:
    """Spatial correlation sampling from input1 to input2 (one fused CUDA kernel instead of P^2 slices).

    Extensions over the reference (both optional): ``input2`` may be a ``WarpedFeatures`` handle (the warp is fused
    into the kernel's operand conversion) and ``_negative_slope`` folds a following LeakyReLU into the epilogue."""
    kernel_size, patch_size, stride = _pair(kernel_size), _pair(patch_size), _pair(stride)
    padding, dilation, dilation_patch = _pair(padding), _pair(dilation), _pair(dilation_patch)
    _validate(kernel_size, patch_size, dilation)
    a = _cabi.require_cuda_f32(input1, "input1")
    flow = None
    if isinstance(input2, WarpedFeatures):
        if warp_fusable(a.shape[1] if a.dim() == 4 else 0):
            b, flow = input2.x, input2.flow
        else:
            b = input2.materialize()
    else:
        b = _cabi.require_cuda_f32(input2, "input2")
    if a.dim() != 4 or a.shape != b.shape:
        raise ValueError(f"input1/input2 must both be (B,C,H,W); got {tuple(input1.shape)} and {tuple(input2.shape)}")
    slope = 1.0 if _negative_slope is None else float(_negative_slope)
    if slope <= 0.0:
        raise ValueError("a fused LeakyReLU needs a positive negative_slope")
    needs_grad = torch.is_grad_enabled() and (a.requires_grad or b.requires_grad or (flow is not None and flow.requires_grad))
    if flow is None and slope == 1.0:
        if needs_grad:
            return _LocalCorrFn.apply(a, b, patch_size, stride, padding, dilation_patch, _scale)
        return local_corr_forward(a, b, patch_size, stride, padding, dilation_patch, _scale)
    if needs_grad:
        return _FusedLocalCorrFn.apply(a, b, flow, patch_size, stride, padding, dilation_patch, _scale, slope)
    return local_corr_forward(a, b, patch_size, stride, padding, dilation_patch, _scale, slope, flow2=flow)


@SIMILARITY_REGISTRY.register()
class IterSpatialCorrelationSampler(nn.Module):
    """
    Spatial correlation sampling from two inputs (FlowNetC, PWC-Net, DCVNet engine), sm_100a kernel.

    Parameters
    ----------
    kernel_size, patch_size, stride, padding, dilation, dilation_patch : int or pair of int
        as in the reference; kernel_size and dilation must be 1, patch_size odd.
    """

    @configurable
    def __init__(
        self,
        kernel_size: IntOrPair = 1,
        patch_size: IntOrPair = 1,
        stride: IntOrPair = 1,
        padding: IntOrPair = 0,
        dilation: IntOrPair = 1,
        dilation_patch: IntOrPair = 1,
    ) -> None:
        super(IterSpatialCorrelationSampler, self).__init__()
        self.kernel_size = kernel_size
        self.patch_size = patch_size
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.dilation_patch = dilation_patch
        # set by ezflow_b200.fuse_model(): slope of the LeakyReLU that followed this sampler, now applied in its epilogue
        self.fused_negative_slope = None

    @classmethod
    def from_config(cls, cfg):
        return {
            "kernel_size": cfg.KERNEL_SIZE,
            "patch_size": cfg.PATCH_SIZE,
            "stride": cfg.STRIDE,
            "padding": cfg.PADDING,
            "dilation": cfg.DILATION,
            "dilation_patch": cfg.DILATION_PATCH,
        }

    def forward(self, input1: torch.Tensor, input2: torch.Tensor) -> torch.Tensor:
        return iter_spatial_correlation_sample(
            input1=input1,
            input2=input2,
            kernel_size=self.kernel_size,
            patch_size=self.patch_size,
            stride=self.stride,
            padding=self.padding,
            dilation=self.dilation,
            dilation_patch=self.dilation_patch,
            _negative_slope=self.fused_negative_slope,
        )
