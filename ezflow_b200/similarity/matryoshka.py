"""DCVNet Matryoshka dilated cost volumes on B200 (K6).

Drop-ins for ``MatryoshkaDilatedCostVolume`` / ``MatryoshkaDilatedCostVolumeList`` (reference
ezflow/similarity/learnable_cost.py:234-434): same constructor arguments, ``from_config`` keys,
buffers (``offsets``, ``offsets_2d``), sub-module names (``cost_volume_list``, ``corr_layers``) and
output layout (B, sum(D)*G, 2r+1, 2r+1, h', w').  All dilations of a scale run in ONE launch that
writes directly into the concatenated output; LeakyReLU(0.1) is fused.
"""

import ctypes

import numpy as np
import torch
import torch.nn as nn

from .. import _cabi
from ..config import configurable
from ..registry import SIMILARITY_REGISTRY
from .sampler import IterSpatialCorrelationSampler as SpatialCorrelationSampler
from .sampler import local_corr_backward, local_corr_workspace


def _matryoshka_forward(x1, x2, G, P, stride, dilations, slope, out=None, ch_offset=0):
    B, C, H, W = x1.shape
    D = len(dilations)
    oh, ow = -(-H // stride), -(-W // stride)
    if out is None:
        out = torch.empty((B, D * G, P, P, oh, ow), dtype=torch.float32, device=x1.device)
    if B == 0:
        return out
    view = out[:, ch_offset : ch_offset + D * G]
    dil = (ctypes.c_int * D)(*[int(d) for d in dilations])
    with torch.cuda.device(x1.device):
        ws, nws = local_corr_workspace(B * G, C // G, H, W, x1.device)
        _cabi.check(
            _cabi.lib().ezb_matryoshka_fwd(
                _cabi.ptr(x1), _cabi.ptr(x2), B, G, C // G, H, W, P, stride, dil, D, float(slope),
                ctypes.c_void_p(view.data_ptr()), int(out.stride(0)), _cabi.ptr(ws), nws, _cabi.stream_ptr(x1.device),
            )
        )
    return out


class _MatryoshkaFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x1, x2, G, P, stride, dilations, slope):
        out = _matryoshka_forward(x1, x2, G, P, stride, dilations, slope)
        ctx.cfg = (G, P, stride, tuple(dilations), slope)
        ctx.save_for_backward(x1, x2, out if slope != 1.0 else None)
        return out

    @staticmethod
    def backward(ctx, dout):
        x1, x2, out = ctx.saved_tensors
        G, P, stride, dilations, slope = ctx.cfg
        dout = dout.contiguous().float()
        if slope != 1.0:
            dout = torch.where(out > 0, dout, dout * slope)  # F.leaky_relu's backward: the slope applies at exactly 0
        B, C, H, W = x1.shape
        x1g = x1.view(B * G, C // G, H, W)
        x2g = x2.view(B * G, C // G, H, W)
        from .sampler import get_local_corr_mode

        lib = _cabi.lib()
        # one call for all dilations (ezb_matryoshka_bwd): tensor-core adjoints accumulate straight into the outputs; the fp32
        # fallback inside it needs one group (DCVNet), otherwise the per-dilation loop below runs
        if G == 1:
            d1 = torch.empty_like(x1g)
            d2 = torch.empty_like(x2g)
            dil = (ctypes.c_int * len(dilations))(*[int(d) for d in dilations])
            n = ctypes.c_size_t(0)
            with torch.cuda.device(x1.device):
                _cabi.check(lib.ezb_local_corr_bwd_workspace_bytes(B * G, C // G, H, W, ctypes.byref(n)))
                ws = torch.empty(n.value, dtype=torch.uint8, device=x1.device)
                _cabi.check(lib.ezb_matryoshka_bwd(_cabi.ptr(x1g), _cabi.ptr(x2g), _cabi.ptr(dout), B, G, C // G, H, W, P, stride, dil, len(dilations),
                                                   int(dout.stride(0)), _cabi.ptr(d1), _cabi.ptr(d2), _cabi.ptr(ws), n.value,
                                                   1 if get_local_corr_mode() == "tc" else 0, _cabi.stream_ptr(x1.device)))
            return d1.view_as(x1), d2.view_as(x2), None, None, None, None, None
        d1 = torch.zeros_like(x1g)
        d2 = torch.zeros_like(x2g)
        for i, d in enumerate(dilations):
            # (B, G, P, P, h, w) slice of dilation i, made contiguous as (B*G, P, P, h, w)
            g = dout[:, i * G : (i + 1) * G].reshape(B * G, P, P, dout.shape[-2], dout.shape[-1]).contiguous()
            a, b = local_corr_backward(x1g, x2g, g, (P, P), (stride, stride), (0, 0), (d, d))
            d1 += a
            d2 += b
        return d1.view_as(x1), d2.view_as(x2), None, None, None, None, None


@SIMILARITY_REGISTRY.register()
class MatryoshkaDilatedCostVolume(nn.Module):
    """
    Cost volume with concentric offset dilations (DCVNet), one fused launch for all dilations.

    Parameters
    ----------
    num_groups : int, default 1
    max_displacement : int, default 4
    stride : int, default 1
    dilations : List[int], default [1, 2, 3, 5, 9, 16]
    use_relu : bool, default False
    """

    @configurable
    def __init__(self, num_groups=1, max_displacement=4, stride=1, dilations=[1, 2, 3, 5, 9, 16], use_relu=False):
        super(MatryoshkaDilatedCostVolume, self).__init__()
        self.num_groups = num_groups
        self.use_relu = use_relu
        self.stride = stride
        self.dilations = [int(d) for d in dilations]
        self.search_range = 2 * max_displacement + 1
        self._set_concentric_offsets(dilations=dilations, radius=max_displacement)
        # parameter-free sub-modules kept so the module tree matches the reference (learnable_cost.py:266-279)
        self.corr_layers = nn.ModuleList(
            [
                SpatialCorrelationSampler(patch_size=self.search_range, stride=stride, padding=0, dilation_patch=d)
                for d in self.dilations
            ]
        )

    @classmethod
    def from_config(cls, cfg):
        return {
            "num_groups": cfg.NUM_GROUPS,
            "max_displacement": cfg.MAX_DISPLACEMENT,
            "stride": cfg.STRIDE,
            "dilations": cfg.DILATIONS,
            "use_relu": cfg.USE_RELU,
        }

    def _set_concentric_offsets(self, dilations, radius):
        offsets = np.array([(np.arange(-radius, radius + 1) * d).tolist() for d in dilations])
        self.register_buffer("offsets", torch.Tensor(offsets).float())

    def get_relative_offsets(self):
        return self.offsets

    def get_search_range(self):
        return self.offsets.shape[1]

    def _forward_into(self, x1, x2, out, ch_offset):
        a = _cabi.require_cuda_f32(x1, "x1")
        b = _cabi.require_cuda_f32(x2, "x2")
        assert a.shape[1] % self.num_groups == 0
        return _matryoshka_forward(
            a, b, self.num_groups, self.search_range, self.stride, self.dilations, 0.1 if self.use_relu else 1.0, out, ch_offset
        )

    def forward(self, x1, x2):
        a = _cabi.require_cuda_f32(x1, "x1")
        b = _cabi.require_cuda_f32(x2, "x2")
        assert a.shape[1] % self.num_groups == 0
        slope = 0.1 if self.use_relu else 1.0
        if torch.is_grad_enabled() and (a.requires_grad or b.requires_grad):
            return _MatryoshkaFn.apply(a, b, self.num_groups, self.search_range, self.stride, self.dilations, slope)
        return _matryoshka_forward(a, b, self.num_groups, self.search_range, self.stride, self.dilations, slope)


def l2_normalize(x, eps=1e-9):
    x = _cabi.require_cuda_f32(x, "x")
    B, C, H, W = x.shape
    out = torch.empty_like(x)
    with torch.cuda.device(x.device):
        _cabi.check(_cabi.lib().ezb_l2_normalize_fwd(_cabi.ptr(x), B, C, H * W, float(eps), _cabi.ptr(out), _cabi.stream_ptr(x.device)))
    return out


@SIMILARITY_REGISTRY.register()
class MatryoshkaDilatedCostVolumeList(nn.Module):
    """
    A list of Matryoshka cost volumes over encoder scales (DCVNet).

    Parameters
    ----------
    num_groups : int, default 1
    max_displacement : int, default 4
    encoder_output_strides : List[int], default [2, 8]
    dilations : List[List[int]], default [[1], [1, 2, 3, 5, 9, 16]]
    normalize_feat_l2 : bool, default False
    use_relu : bool, default False
    """

    @configurable
    def __init__(
        self,
        num_groups=1,
        max_displacement=4,
        encoder_output_strides=[2, 8],
        dilations=[[1], [1, 2, 3, 5, 9, 16]],
        normalize_feat_l2=False,
        use_relu=False,
    ):
        super(MatryoshkaDilatedCostVolumeList, self).__init__()
        self.normalize_feat_l2 = normalize_feat_l2
        self.cost_volume_list = nn.ModuleList()
        offsets = None
        for dilations_i, feat_stride_i in zip(dilations, encoder_output_strides):
            assert feat_stride_i <= 8
            cv = MatryoshkaDilatedCostVolume(
                num_groups=num_groups,
                max_displacement=max_displacement,
                dilations=dilations_i,
                stride=8 // feat_stride_i,
                use_relu=use_relu,
            )
            self.cost_volume_list.append(cv)
            rel = cv.get_relative_offsets() * feat_stride_i
            offsets = rel if offsets is None else torch.cat((offsets, rel), dim=0)
        self.offsets = offsets
        self._set_global_flow_offsets()

    @classmethod
    def from_config(cls, cfg):
        return {
            "num_groups": cfg.NUM_GROUPS,
            "max_displacement": cfg.MAX_DISPLACEMENT,
            "encoder_output_strides": cfg.ENCODER_OUTPUT_STRIDES,
            "dilations": cfg.DILATIONS,
            "normalize_feat_l2": cfg.NORMALIZE_FEAT_L2,
            "use_relu": cfg.USE_RELU,
        }

    def _set_global_flow_offsets(self):
        num_dilations, search_range = self.offsets.shape
        offsets_2d = torch.zeros((num_dilations, search_range, search_range, 2))
        for idx in range(num_dilations):
            oi, oj = torch.meshgrid(self.offsets[idx], self.offsets[idx], indexing="ij")
            offsets_2d[idx, :, :, 0] = oi  # y
            offsets_2d[idx, :, :, 1] = oj  # x
        self.register_buffer("offsets_2d", torch.Tensor(offsets_2d).float())

    def get_global_flow_offsets(self):
        return self.offsets_2d

    def get_search_range(self):
        return self.cost_volume_list[0].get_search_range()

    def forward(self, x1, x2):
        needs_grad = torch.is_grad_enabled() and any(t.requires_grad for t in list(x1) + list(x2))
        if needs_grad:
            costs = []
            for idx in range(len(x1)):
                a, b = x1[idx], x2[idx]
                if self.normalize_feat_l2:
                    a = a / (a.norm(dim=1, keepdim=True) + 1e-9)
                    b = b / (b.norm(dim=1, keepdim=True) + 1e-9)
                costs.append(self.cost_volume_list[idx](a, b))
            return torch.cat(costs, dim=1)
        # inference: every scale writes straight into its channel slice of the concatenated volume
        out = None
        ch = 0
        total_ch = sum(len(cv.dilations) * cv.num_groups for cv in self.cost_volume_list)
        for idx, cv in enumerate(self.cost_volume_list):
            a = _cabi.require_cuda_f32(x1[idx], f"x1[{idx}]")
            b = _cabi.require_cuda_f32(x2[idx], f"x2[{idx}]")
            if self.normalize_feat_l2:
                a, b = l2_normalize(a), l2_normalize(b)
            B, _, H, W = a.shape
            oh, ow = -(-H // cv.stride), -(-W // cv.stride)
            if out is None:
                P = cv.search_range
                out = torch.empty((B, total_ch, P, P, oh, ow), dtype=torch.float32, device=a.device)
            elif (oh, ow) != tuple(out.shape[-2:]):
                raise RuntimeError(
                    f"Sizes of tensors must match except in dimension 1: scale {idx} gives {(oh, ow)}, expected {tuple(out.shape[-2:])}"
                )
            cv._forward_into(a, b, out, ch)
            ch += len(cv.dilations) * cv.num_groups
        return out
