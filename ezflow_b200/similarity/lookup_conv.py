"""Lookup -> 1x1 convolution fusion for RAFT's update block.

The reference's motion encoders start with ``cor = F.relu(self.convc1(corr))`` on the lookup features
(ezflow/decoder/iterative/recurrent_lookup.py:153,182,204), ``convc1 = Conv2d(L*(2r+1)^2, 256 or 96, 1)``.  With
``fuse_model(model)`` the correlation object hands the update block a ``LazyLookup`` handle instead of the
(B, L*(2r+1)^2, H, W) tensor, and ``convc1`` (re-classed to ``LookupConv2d``, same parameters and state-dict keys) runs
the lookup, the 1x1 convolution, the bias and the ReLU in one tcgen05 kernel (csrc/corr_lookup_conv.cu): the lookup
features never reach HBM.  Anything the fused kernel does not cover (gradients, fp32 pyramids, > 256 output channels)
materialises the handle and runs the two separate ops.
"""

import ctypes

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import _cabi

FUSED_MAX_RADIUS = 4
FUSED_MAX_OUT = 256


class LazyLookup:
    """Deferred ``corr_fn(coords)`` (reference pairwise.py:43-69).  ``materialize()`` gives the ordinary tensor."""

    def __init__(self, corr_obj, coords):
        self.corr_obj = corr_obj
        self.coords = coords
        st = corr_obj._state
        k = 2 * int(corr_obj.corr_radius) + 1
        self._shape = (st.B, st.L * k * k, st.H, st.W)

    shape = property(lambda self: torch.Size(self._shape))
    device = property(lambda self: self.coords.device)
    dtype = property(lambda self: torch.float32)

    def dim(self):
        return 4

    def size(self, i=None):
        return self.shape if i is None else self._shape[i]

    def materialize(self):
        return self.corr_obj._lookup(self.coords)


def _packed_weights(weight, L, radius):
    """fp16 K-major copy of the 1x1 weights in the kernel's layout + the power-of-two factor that undoes its scaling.
    Cached ON the tensor object (a module's Parameter lives as long as the module) and invalidated by its version counter;
    a cache keyed by data_ptr would go stale when the allocator hands the same address to another tensor."""
    key = (weight._version, L, radius, weight.data_ptr())
    hit = getattr(weight, "_ezb_packed", None)
    if hit is not None and hit[0] == key:
        return hit[1], hit[2]
    lib = _cabi.lib()
    n = ctypes.c_size_t(0)
    _cabi.check(lib.ezb_lookup_conv_packed_weight_bytes(L, ctypes.byref(n)))
    packed = torch.empty(n.value, dtype=torch.uint8, device=weight.device)
    wscale = torch.empty(1, dtype=torch.float32, device=weight.device)
    w = weight.detach().reshape(weight.shape[0], -1).float().contiguous()
    with torch.cuda.device(weight.device):
        _cabi.check(lib.ezb_lookup_conv_pack_weights(_cabi.ptr(w), w.shape[0], L, radius, _cabi.ptr(packed), _cabi.ptr(wscale),
                                                     _cabi.stream_ptr(weight.device)))
    try:
        weight._ezb_packed = (key, packed, wscale)
    except AttributeError:
        pass
    return packed, wscale


def fusable(lazy, weight):
    st = lazy.corr_obj._state
    k = 2 * int(lazy.corr_obj.corr_radius) + 1
    return (
        st.dtype_code == _cabi.EZB_F16
        and int(lazy.corr_obj.corr_radius) <= FUSED_MAX_RADIUS
        and weight.dim() == 4 and weight.shape[2] == 1 and weight.shape[3] == 1
        and weight.shape[1] == st.L * k * k
        and weight.shape[0] <= FUSED_MAX_OUT
        and weight.is_cuda and weight.device == st.device
        and not (torch.is_grad_enabled() and weight.requires_grad)
    )


def lookup_conv1x1(lazy, weight, bias=None, relu=True):
    """relu(conv1x1(lookup(coords), weight, bias)) as one kernel; (B, Cout, H, W) fp32."""
    if not fusable(lazy, weight):
        out = F.conv2d(lazy.materialize(), weight, bias)
        return F.relu(out) if relu else out
    st = lazy.corr_obj._state
    radius = int(lazy.corr_obj.corr_radius)
    packed, wscale = _packed_weights(weight, st.L, radius)
    cout = weight.shape[0]
    out = torch.empty((st.B, cout, st.H, st.W), dtype=torch.float32, device=st.device)
    b = bias.detach().float().contiguous() if bias is not None else None
    with torch.cuda.device(st.device):
        _cabi.check(
            _cabi.lib().ezb_corr_lookup_conv1x1_fwd(
                _cabi.ptr(st.pyr), _cabi.ptr(st.deq), _cabi.ptr(lazy.coords), st.B, st.H, st.W, st.L, radius, st.dtype_code,
                _cabi.ptr(packed), _cabi.ptr(wscale), _cabi.ptr(b), cout, 1 if relu else 0, _cabi.ptr(out),
                _cabi.stream_ptr(st.device),
            )
        )
    return out


class LookupConv2d(nn.Conv2d):
    """``nn.Conv2d`` (1x1) that also accepts a ``LazyLookup``: then lookup + convolution + bias + ReLU run fused and the
    result is marked as already activated.  An existing conv is re-classed in place (``fuse_model``), so parameters,
    hooks and state-dict keys stay exactly the reference's."""

    def forward(self, x):
        if isinstance(x, LazyLookup):
            out = lookup_conv1x1(x, self.weight, self.bias, relu=True)
            out._ezb_relu_applied = True
            return out
        return super().forward(x)


def fused_motion_encoder_forward(self, flow, corr):
    """MotionEncoder.forward / SmallMotionEncoder.forward (reference ezflow/decoder/iterative/recurrent_lookup.py:139-160,
    190-212) with the first two ops — convc1 and its ReLU — taken by the fused kernel when ``corr`` is a LazyLookup."""
    cor = self.convc1(corr)
    if not getattr(cor, "_ezb_relu_applied", False):
        cor = F.relu(cor)
    if hasattr(self, "convc2"):
        cor = F.relu(self.convc2(cor))
    flo = F.relu(self.convf1(flow))
    flo = F.relu(self.convf2(flo))
    out = F.relu(self.conv(torch.cat([cor, flo], dim=1)))
    return torch.cat([out, flow], dim=1)
