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
        shift = self.pad_size - self.max_h_disp
        if shift > 0:
            # reference layer.py:44-60: offsets 0..2d into features2 padded by pad_size, i.e. displacements -pad..2d-pad —
            # an off-centre window.  On canvases grown by `shift` (features1 at the top-left, features2 at the bottom-right)
            # it is the centred window; the extra output rows / columns are cropped below.
            features1 = nn.functional.pad(features1, (0, shift, 0, shift))
            features2 = nn.functional.pad(features2, (shift, 0, shift, 0))
        p = 2 * self.max_h_disp + 1
        c = features1.shape[1]
        out = iter_spatial_correlation_sample(features1, features2, patch_size=p, _scale=1.0 / c)
        b, _, _, h, w = out.shape
        out = out.view(b, p * p, h, w)
        return out[:, :, : h - shift, : w - shift].contiguous() if shift > 0 else out
