"""GPU parity: convex flow upsampling (K7; reference ezflow/utils/resampling.py:112-141).

fp32 throughout; the only differences from the reference are summation order and expf rounding: 2e-5 relative."""

import numpy as np
import pytest
import torch

from gpu_util import assert_close, dev, to_gpu
from oracle import similarity_oracle as O

pytestmark = pytest.mark.gpu

TOL = 2e-5


@pytest.mark.parametrize("name", ["convex_up_s8", "convex_up_s4"])
def test_convex_upsample_vs_reference_golden(golden, name):
    import ezflow_b200 as E

    g = golden(name)
    s = int(g["out_stride"])
    flow, logits = to_gpu(g["flow"], True), to_gpu(g["logits"], True)
    out = E.convex_upsample_flow(flow, logits, s)
    assert out.dtype == torch.float32 and out.is_contiguous()
    assert_close(out, g["out"], TOL, "convex upsample fwd")
    (out * to_gpu(g["dout"])).sum().backward()
    assert_close(flow.grad, g["dflow"], TOL, "convex upsample dflow")
    assert_close(logits.grad, g["dlogits"], 5e-5, "convex upsample dlogits")


@pytest.mark.parametrize(
    "shape",
    [
        (1, 2, 46, 62, 8),  # BASELINE config 1 (368x496 / 8)
        (2, 2, 1, 1, 8),  # a single coarse pixel: every neighbour is padding
        (1, 1, 3, 70, 2),  # ragged width (3 tiles of 32), one channel
        (1, 4, 5, 33, 3),  # out_stride that is not a power of two, 4 channels
        (1, 2, 7, 31, 16),  # large out_stride (opt-in shared memory)
    ],
)
def test_convex_upsample_vs_oracle_shapes(shape):
    import ezflow_b200 as E

    n, c, h, w, s = shape
    rng = np.random.default_rng(sum(shape))
    flow = (rng.standard_normal((n, c, h, w)) * 5).astype(np.float32)
    logits = (rng.standard_normal((n, 9 * s * s, h, w)) * 3).astype(np.float32)
    dout = rng.standard_normal((n, c, s * h, s * w)).astype(np.float32)
    fg, lg = to_gpu(flow, True), to_gpu(logits, True)
    out = E.convex_upsample_flow(fg, lg, s)
    assert_close(out, O.convex_upsample_flow(flow, logits, s), TOL, f"fwd {shape}")
    (out * to_gpu(dout)).sum().backward()
    dflow, dlogits = O.convex_upsample_flow_bwd(flow, logits, s, dout)
    assert_close(fg.grad, dflow, TOL, f"dflow {shape}")
    assert_close(lg.grad, dlogits, 5e-5, f"dlogits {shape}")
    # only one gradient requested
    fg2 = to_gpu(flow, True)
    (E.convex_upsample_flow(fg2, to_gpu(logits), s) * to_gpu(dout)).sum().backward()
    assert_close(fg2.grad, dflow, TOL, f"dflow only {shape}")
    lg2 = to_gpu(logits, True)
    (E.convex_upsample_flow(to_gpu(flow), lg2, s) * to_gpu(dout)).sum().backward()
    assert_close(lg2.grad, dlogits, 5e-5, f"dlogits only {shape}")


def test_convex_upsample_1080p_properties():
    """Headline size (135x240 -> 1080x1920): properties that need no oracle."""
    import ezflow_b200 as E

    g = torch.Generator(device=dev()).manual_seed(7)
    n, h, w, s = 2, 135, 240, 8
    flow = torch.randn(n, 2, h, w, device=dev(), generator=g) * 8
    logits = torch.randn(n, 9 * s * s, h, w, device=dev(), generator=g)
    out = E.convex_upsample_flow(flow, logits, s)
    assert tuple(out.shape) == (n, 2, 1080, 1920)
    # convex combination: every fine value lies inside the hull of its 3x3 coarse neighbourhood (with the zero padding)
    pad = torch.nn.functional.pad(flow, (1, 1, 1, 1))
    nb = torch.stack([pad[:, :, ky : ky + h, kx : kx + w] for ky in range(3) for kx in range(3)], 0)
    lo = nb.min(0).values.repeat_interleave(s, 2).repeat_interleave(s, 3)
    hi = nb.max(0).values.repeat_interleave(s, 2).repeat_interleave(s, 3)
    assert bool(((out >= lo - 1e-4) & (out <= hi + 1e-4)).all())
    # a constant flow stays constant away from the border; logits shifted per pixel do not change the result
    const = torch.full_like(flow, 3.25)
    up = E.convex_upsample_flow(const, logits, s)
    assert torch.allclose(up[:, :, s:-s, s:-s], torch.full_like(up[:, :, s:-s, s:-s], 3.25), atol=1e-5)
    shift = torch.randn(n, 1, h, w, device=dev(), generator=g) * 4
    assert torch.allclose(E.convex_upsample_flow(flow, logits + shift, s), out, atol=2e-4)
    # one-hot logits select a single neighbour: tap 4 is the centre -> nearest-neighbour upsampling
    hot = torch.full_like(logits, -1e4).view(n, 9, s * s, h, w)
    hot[:, 4] = 0
    nn_up = E.convex_upsample_flow(flow, hot.view_as(logits), s)
    assert torch.equal(nn_up, flow.repeat_interleave(s, 2).repeat_interleave(s, 3))
    # torch fp32 restatement at full size
    probs = torch.softmax(logits.view(n, 1, 9, s, s, h, w), dim=2)
    ref = (probs * nb.permute(1, 2, 0, 3, 4).reshape(n, 2, 9, 1, 1, h, w)).sum(2).permute(0, 1, 4, 2, 5, 3).reshape(n, 2, s * h, s * w)
    assert_close(out, ref, TOL, "1080p vs torch fp32")
