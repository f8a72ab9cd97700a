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


def local_corr_workspace(BG, C, H, W, device, padding=(0, 0)):
    """Scratch for the tensor-core path (fp16 K-major operand copies + scales), or (None, 0) in 'exact' mode."""
    if _LOCAL_CORR_MODE != "tc":
        return None, 0
    n = ctypes.c_size_t(0)
    _cabi.check(_cabi.lib().ezb_local_corr_workspace_bytes(BG, C, H, W, int(padding[0]), int(padding[1]), ctypes.byref(n)))
    return torch.empty(n.value, dtype=torch.uint8, device=device), n.value


def local_corr_forward(in1, in2, patch, stride, padding, dpatch, scale=1.0, slope=1.0, out=None, out_batch_stride=0):
    B, C, H, W = in1.shape
    oh, ow = _out_hw(H, W, stride, padding)
    if out is None:
        out = torch.empty((B, patch[0], patch[1], oh, ow), dtype=torch.float32, device=in1.device)
    with torch.cuda.device(in1.device):
        ws, nws = local_corr_workspace(B, C, H, W, in1.device, padding)
        _cabi.check(
            _cabi.lib().ezb_local_corr_fwd(
                _cabi.ptr(in1), _cabi.ptr(in2), B, C, H, W, patch[0], patch[1], stride[0], stride[1], padding[0], padding[1],
                dpatch[0], dpatch[1], float(scale), float(slope), _cabi.ptr(out), int(out_batch_stride),
                _cabi.ptr(ws), nws, _cabi.stream_ptr(in1.device),
            )
        )
    return out


def local_corr_backward(in1, in2, dout, patch, stride, padding, dpatch, scale=1.0, need1=True, need2=True, dout_batch_stride=0):
    B, C, H, W = in1.shape
    d1 = torch.empty_like(in1) if need1 else None
    d2 = torch.empty_like(in2) if need2 else None
    with torch.cuda.device(in1.device):
        _cabi.check(
            _cabi.lib().ezb_local_corr_bwd(
                _cabi.ptr(in1), _cabi.ptr(in2), _cabi.ptr(dout), B, C, H, W, patch[0], patch[1], stride[0], stride[1],
                padding[0], padding[1], dpatch[0], dpatch[1], float(scale), int(dout_batch_stride), _cabi.ptr(d1), _cabi.ptr(d2),
                _cabi.stream_ptr(in1.device),
            )
        )
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
) -> torch.Tensor:
    """Spatial correlation sampling from input1 to input2 (one fused CUDA kernel instead of P^2 slices)."""
    kernel_size, patch_size, stride = _pair(kernel_size), _pair(patch_size), _pair(stride)
    padding, dilation, dilation_patch = _pair(padding), _pair(dilation), _pair(dilation_patch)
    _validate(kernel_size, patch_size, dilation)
    a = _cabi.require_cuda_f32(input1, "input1")
    b = _cabi.require_cuda_f32(input2, "input2")
    if a.dim() != 4 or a.shape != b.shape:
        raise ValueError(f"input1/input2 must both be (B,C,H,W); got {tuple(input1.shape)} and {tuple(input2.shape)}")
    if torch.is_grad_enabled() and (a.requires_grad or b.requires_grad):
        return _LocalCorrFn.apply(a, b, patch_size, stride, padding, dilation_patch, _scale)
    return local_corr_forward(a, b, patch_size, stride, padding, dilation_patch, _scale)


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
        )
