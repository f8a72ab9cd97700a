"""CorrelationLayer on B200: K4 with a 1/C epilogue.

Drop-in for ``ezflow.similarity.correlation.layer.CorrelationLayer`` (reference
ezflow/similarity/correlation/layer.py:11-67): (2d+1)^2 window, *mean* over channels, output
(B, (2d+1)^2, H, W) with the vertical offset as the slow channel index.
"""

import torch.nn as nn

from ..config import configurable
from ..registry import SIMILARITY_REGISTRY
from .sampler import iter_spatial_correlation_sample


@SIMILARITY_REGISTRY.register()
class CorrelationLayer(nn.Module):
    """
    Parameters
    ----------
    pad_size : int
        Padding size for the correlation layer
    max_displacement : int
        Maximum displacement for the correlation computation
    """

    @configurable
    def __init__(self, pad_size=4, max_displacement=4):
        super().__init__()
        self.max_h_disp = max_displacement
        self.pad_size = pad_size
        self.padlayer = nn.ConstantPad2d(pad_size, 0)  # kept for attribute compatibility; padding is implicit in the kernel

    @classmethod
    def from_config(cls, cfg):
        return {
            "pad_size": cfg.PAD_SIZE,
            "max_displacement": cfg.MAX_DISPLACEMENT,
        }

    def forward(self, features1, features2):
        if self.pad_size < self.max_h_disp:
            # the reference's slices of features2 padded by pad_size run out of range: it cannot run this either
            raise NotImplementedError("CorrelationLayer needs pad_size >= max_displacement (as the reference does).")
        if self.pad_size > self.max_h_disp:
            # reference layer.py:44-60: offsets 0..2d into features2 padded by pad_size, i.e. displacements -pad..2d-pad —
            # an off-centre window.  Shifting features2 by (pad - d) pixels (zeros shifted in) turns it into the centred one.
            s = self.pad_size - self.max_h_disp
            H, W = features2.shape[-2:]
            features2 = nn.functional.pad(features2, (s, 0, s, 0))[:, :, :H, :W]
        p = 2 * self.max_h_disp + 1
        c = features1.shape[1]
        out = iter_spatial_correlation_sample(features1, features2, patch_size=p, _scale=1.0 / c)
        b, _, _, h, w = out.shape
        return out.view(b, p * p, h, w)
