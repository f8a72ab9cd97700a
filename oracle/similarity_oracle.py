"""CPU oracle for the ezflow similarity / cost-volume hot path (numpy).

TEST INFRASTRUCTURE ONLY.  This module is a plain-numpy restatement of the
reference's (neu-vi/ezflow v0.2.5) PyTorch arithmetic for the path named in
BASELINE.json.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the
product package ``ezflow_b200`` never does (it fails loudly when the CUDA
library is missing instead of falling back to anything here).

Parity pinning: the reference's own tests hold no value-level golden vectors
for this path, so the oracle is pinned against outputs of the
reference itself: ``tools/make_golden.py`` imports the unmodified reference from
``/root/reference`` (through the fvcore/omegaconf shim in ``tests/refshim.py``),
runs its classes on seeded inputs and commits the results under
``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` checks every function
below against those vectors.

Each function cites the reference file:line (relative to /root/reference) it
follows.  All arithmetic is fp32 unless ``dtype=np.float64`` is requested.
"""

from __future__ import annotations

import math
from typing import List, Sequence, Tuple, Union

import numpy as np

IntOrPair = Union[int, Tuple[int, int]]


def _pair(v: IntOrPair) -> Tuple[int, int]:
    return (v, v) if isinstance(v, int) else (int(v[0]), int(v[1]))


# --------------------------------------------------------------------------
# K1: all-pairs correlation       ezflow/similarity/correlation/pairwise.py:78-88
# --------------------------------------------------------------------------
def corr_all_pairs(fmap1: np.ndarray, fmap2: np.ndarray, dtype=np.float32) -> np.ndarray:
    """corr[b,h1,w1,0,h2,w2] = sum_c f1[b,c,h1,w1] f2[b,c,h2,w2] / sqrt(C)."""
    b, c, h, w = fmap1.shape
    f1 = fmap1.reshape(b, c, h * w).astype(dtype)
    f2 = fmap2.reshape(b, c, h * w).astype(dtype)
    corr = np.matmul(f1.transpose(0, 2, 1), f2)
    corr = corr.reshape(b, h, w, 1, h, w)
    return (corr / dtype(math.sqrt(float(c)))).astype(dtype)


# --------------------------------------------------------------------------
# K2: average-pool pyramid        ezflow/similarity/correlation/pairwise.py:26-41
# --------------------------------------------------------------------------
def avg_pool2x2(x: np.ndarray) -> np.ndarray:
    """F.avg_pool2d(x, 2, stride=2) with the default ceil_mode=False (floor)."""
    h, w = x.shape[-2:]
    ho, wo = h // 2, w // 2
    x = x[..., : 2 * ho, : 2 * wo]
    x = x.reshape(*x.shape[:-2], ho, 2, wo, 2)
    # ATen sums the window then divides by the window size
    return (x.sum(axis=(-3, -1)) / x.dtype.type(4.0)).astype(x.dtype)


def corr_pyramid(fmap1, fmap2, num_levels: int = 4, dtype=np.float32) -> List[np.ndarray]:
    corr = corr_all_pairs(fmap1, fmap2, dtype)
    b, h1, w1, dim, h2, w2 = corr.shape
    corr = corr.reshape(b * h1 * w1, dim, h2, w2)
    pyr = [corr]
    for _ in range(num_levels - 1):
        corr = avg_pool2x2(corr)
        pyr.append(corr)
    return pyr


# --------------------------------------------------------------------------
# grid_sample(bilinear, zeros, align_corners=True)   (ATen GridSampler semantics)
# used by ezflow/utils/resampling.py:56-85 and ezflow/utils/warp.py:34-37
# --------------------------------------------------------------------------
def grid_sample_bilinear_zeros_ac(img: np.ndarray, grid: np.ndarray) -> np.ndarray:
    """img (N,C,H,W), grid (N,Ho,Wo,2) normalised [-1,1] (x,y) -> (N,C,Ho,Wo)."""
    n, c, h, w = img.shape
    dt = img.dtype.type
    gx = grid[..., 0].astype(img.dtype)
    gy = grid[..., 1].astype(img.dtype)
    ix = ((gx + dt(1)) / dt(2)) * dt(w - 1)
    iy = ((gy + dt(1)) / dt(2)) * dt(h - 1)
    x0 = np.floor(ix)
    y0 = np.floor(iy)
    wx1 = ix - x0
    wx0 = dt(1) - wx1
    wy1 = iy - y0
    wy0 = dt(1) - wy1
    x0i = x0.astype(np.int64)
    y0i = y0.astype(np.int64)
    out = np.zeros((n, c) + gx.shape[1:], dtype=img.dtype)
    nidx = np.arange(n).reshape(n, 1, 1)
    for dy, wy in ((0, wy0), (1, wy1)):
        for dx, wx in ((0, wx0), (1, wx1)):
            xi = x0i + dx
            yi = y0i + dy
            valid = (xi >= 0) & (xi < w) & (yi >= 0) & (yi < h)
            xc = np.clip(xi, 0, w - 1)
            yc = np.clip(yi, 0, h - 1)
            v = img[nidx, :, yc, xc]  # (N,Ho,Wo,C)
            wgt = (wy * wx * valid.astype(img.dtype))[..., None]
            out += np.moveaxis(v * wgt, -1, 1)
    return out


def bilinear_sampler(img: np.ndarray, coords: np.ndarray) -> np.ndarray:
    """ezflow/utils/resampling.py:56-85 (pixel coords -> normalised -> grid_sample)."""
    h, w = img.shape[-2:]
    dt = img.dtype.type
    x = coords[..., 0:1].astype(img.dtype)
    y = coords[..., 1:2].astype(img.dtype)
    with np.errstate(divide="ignore", invalid="ignore"):
        xg = dt(2) * x / dt(w - 1) - dt(1)
        yg = dt(2) * y / dt(h - 1) - dt(1)
    return grid_sample_bilinear_zeros_ac(img, np.concatenate([xg, yg], axis=-1))


# --------------------------------------------------------------------------
# K3: pyramid lookup              ezflow/similarity/correlation/pairwise.py:43-69
# --------------------------------------------------------------------------
def corr_lookup(pyramid: Sequence[np.ndarray], coords: np.ndarray, radius: int = 4) -> np.ndarray:
    """coords (B,2,H,W) (ch0=x, ch1=y) -> (B, L*(2r+1)^2, H, W) fp32.

    Output channel l*(2r+1)^2 + i*(2r+1) + j samples level l at
    (x/2^l + (i-r), y/2^l + (j-r)): the *slow* window index offsets x."""
    r = radius
    dtype = pyramid[0].dtype
    c = np.transpose(coords, (0, 2, 3, 1)).astype(dtype)
    b, h1, w1, _ = c.shape
    d = np.linspace(-r, r, 2 * r + 1).astype(dtype)
    # delta[i,j] = (d[i], d[j]) and is *added to (x,y)*         pairwise.py:53-61
    delta = np.stack(np.meshgrid(d, d, indexing="ij"), axis=-1)
    outs = []
    for lvl, corr in enumerate(pyramid):
        cen = c.reshape(b * h1 * w1, 1, 1, 2) / dtype.type(2**lvl)
        coords_lvl = cen + delta.reshape(1, 2 * r + 1, 2 * r + 1, 2)
        s = bilinear_sampler(corr, coords_lvl)  # (BN,1,2r+1,2r+1)
        outs.append(s.reshape(b, h1, w1, -1))
    out = np.concatenate(outs, axis=-1)
    return np.ascontiguousarray(np.transpose(out, (0, 3, 1, 2))).astype(np.float32)


def corr_lookup_bwd(
    pyramid_shapes: Sequence[Tuple[int, int]], coords: np.ndarray, dout: np.ndarray, radius: int = 4
) -> List[np.ndarray]:
    """Adjoint of corr_lookup w.r.t. the pyramid (autograd of grid_sample, zeros padding).

    pyramid_shapes[l] = (h_l, w_l); returns list of (B*N,1,h_l,w_l) gradients.
    The gradient w.r.t. coords is corr_lookup_bwd_coords below."""
    r = radius
    k = 2 * r + 1
    b, _, h1, w1 = coords.shape
    n = b * h1 * w1
    dt = np.float32
    c = np.transpose(coords, (0, 2, 3, 1)).reshape(n, 2).astype(dt)
    do = np.transpose(dout, (0, 2, 3, 1)).reshape(n, len(pyramid_shapes), k, k).astype(dt)
    grads = []
    d = np.arange(-r, r + 1).astype(dt)
    for lvl, (hl, wl) in enumerate(pyramid_shapes):
        g = np.zeros((n, hl, wl), dtype=dt)
        cx = c[:, 0] / dt(2**lvl)
        cy = c[:, 1] / dt(2**lvl)
        sx = cx[:, None, None] + d[None, :, None]  # (n,i,1)
        sy = cy[:, None, None] + d[None, None, :]  # (n,1,j)
        with np.errstate(divide="ignore", invalid="ignore"):
            gx = dt(2) * sx / dt(wl - 1) - dt(1)
            gy = dt(2) * sy / dt(hl - 1) - dt(1)
        ix = np.broadcast_to(((gx + dt(1)) / dt(2)) * dt(wl - 1), (n, k, k))
        iy = np.broadcast_to(((gy + dt(1)) / dt(2)) * dt(hl - 1), (n, k, k))
        x0 = np.floor(ix)
        y0 = np.floor(iy)
        wx1 = ix - x0
        wy1 = iy - y0
        qidx = np.broadcast_to(np.arange(n)[:, None, None], (n, k, k))
        for dy, wy in ((0, dt(1) - wy1), (1, wy1)):
            for dx, wx in ((0, dt(1) - wx1), (1, wx1)):
                xi = x0.astype(np.int64) + dx
                yi = y0.astype(np.int64) + dy
                valid = (xi >= 0) & (xi < wl) & (yi >= 0) & (yi < hl)
                np.add.at(
                    g,
                    (qidx[valid], yi[valid], xi[valid]),
                    (do[:, lvl] * wy * wx)[valid],
                )
        grads.append(g.reshape(n, 1, hl, wl))
    return grads


def corr_lookup_bwd_coords(pyramid: Sequence[np.ndarray], coords: np.ndarray, dout: np.ndarray, radius: int = 4) -> np.ndarray:
    """Adjoint of corr_lookup w.r.t. coords -> (B,2,H,W): autograd of ``coords / 2**l`` (pairwise.py:50-52), the
    normalisation in bilinear_sampler (utils/resampling.py:71-77) and grid_sample's grid gradient (bilinear, zeros
    padding, align_corners=True).  RAFT never uses it (``coords1.detach()``, models/raft.py:115); the reference's op
    produces it all the same.

    d sample / d ix = (1-fy)*(v01 - v00) + fy*(v11 - v10), d sample / d iy = (1-fx)*(v10 - v00) + fx*(v11 - v01), where
    neighbours outside the level count as 0; the (W_l-1)/2 of the un-normalisation and the 2/(W_l-1) of the
    normalisation cancel up to rounding, leaving the factor 1/2**l per level."""
    r = radius
    k = 2 * r + 1
    b, _, h1, w1 = coords.shape
    n = b * h1 * w1
    dt = np.float32
    c = np.transpose(coords, (0, 2, 3, 1)).reshape(n, 2).astype(dt)
    do = np.transpose(dout, (0, 2, 3, 1)).reshape(n, len(pyramid), k, k).astype(dt)
    d = np.arange(-r, r + 1).astype(dt)
    g = np.zeros((n, 2), dtype=np.float64)
    qidx = np.broadcast_to(np.arange(n)[:, None, None], (n, k, k))
    for lvl, corr in enumerate(pyramid):
        hl, wl = corr.shape[-2:]
        img = corr.reshape(n, hl, wl).astype(dt)
        sx = c[:, 0, None, None] / dt(2**lvl) + d[None, :, None]
        sy = c[:, 1, None, None] / dt(2**lvl) + d[None, None, :]
        with np.errstate(divide="ignore", invalid="ignore"):
            gx = dt(2) * sx / dt(wl - 1) - dt(1)
            gy = dt(2) * sy / dt(hl - 1) - dt(1)
        ix = np.broadcast_to(((gx + dt(1)) / dt(2)) * dt(wl - 1), (n, k, k))
        iy = np.broadcast_to(((gy + dt(1)) / dt(2)) * dt(hl - 1), (n, k, k))
        x0 = np.floor(ix)
        y0 = np.floor(iy)
        fx = ix - x0
        fy = iy - y0
        v = {}
        for dy in (0, 1):
            for dx in (0, 1):
                xi = x0.astype(np.int64) + dx
                yi = y0.astype(np.int64) + dy
                valid = (xi >= 0) & (xi < wl) & (yi >= 0) & (yi < hl)
                v[dy, dx] = np.where(valid, img[qidx, np.clip(yi, 0, hl - 1), np.clip(xi, 0, wl - 1)], dt(0))
        dix = (dt(1) - fy) * (v[0, 1] - v[0, 0]) + fy * (v[1, 1] - v[1, 0])
        diy = (dt(1) - fx) * (v[1, 0] - v[0, 0]) + fx * (v[1, 1] - v[0, 1])
        g[:, 0] += (do[:, lvl] * dix).sum(axis=(1, 2), dtype=np.float64) / 2**lvl
        g[:, 1] += (do[:, lvl] * diy).sum(axis=(1, 2), dtype=np.float64) / 2**lvl
    return np.ascontiguousarray(np.transpose(g.reshape(b, h1, w1, 2), (0, 3, 1, 2))).astype(np.float32)


def avg_pool2x2_bwd(dy: np.ndarray, in_hw: Tuple[int, int]) -> np.ndarray:
    """Adjoint of floor avg-pool: each of the 4 inputs gets dy/4; dropped row/col get 0."""
    h, w = in_hw
    ho, wo = h // 2, w // 2
    dx = np.zeros(dy.shape[:-2] + (h, w), dtype=dy.dtype)
    up = np.repeat(np.repeat(dy, 2, axis=-2), 2, axis=-1) / dy.dtype.type(4)
    dx[..., : 2 * ho, : 2 * wo] = up
    return dx


def corr_pyramid_bwd(fmap1, fmap2, dpyr: Sequence[np.ndarray]):
    """Adjoint of corr_pyramid: fold level grads into level 0, then the two GEMMs."""
    b, c, h, w = fmap1.shape
    n = h * w
    g = dpyr[-1].astype(np.float32)
    for lvl in range(len(dpyr) - 2, -1, -1):
        g = dpyr[lvl].astype(np.float32) + avg_pool2x2_bwd(g, dpyr[lvl].shape[-2:])
    g = g.reshape(b, n, n) / np.float32(math.sqrt(float(c)))
    f1 = fmap1.reshape(b, c, n).astype(np.float32)
    f2 = fmap2.reshape(b, c, n).astype(np.float32)
    df1 = np.matmul(f2, g.transpose(0, 2, 1)).reshape(b, c, h, w)  # df1[c,q] = sum_t g[q,t] f2[c,t]
    df2 = np.matmul(f1, g).reshape(b, c, h, w)  # df2[c,t] = sum_q g[q,t] f1[c,q]
    return df1, df2


# --------------------------------------------------------------------------
# K4: local displacement-window correlation   similarity/correlation/sampler.py:29-129
# --------------------------------------------------------------------------
def iter_spatial_correlation_sample(
    input1: np.ndarray,
    input2: np.ndarray,
    kernel_size: IntOrPair = 1,
    patch_size: IntOrPair = 1,
    stride: IntOrPair = 1,
    padding: IntOrPair = 0,
    dilation: IntOrPair = 1,
    dilation_patch: IntOrPair = 1,
) -> np.ndarray:
    kernel_size, patch_size = _pair(kernel_size), _pair(patch_size)
    stride, padding = _pair(stride), _pair(padding)
    dilation, dilation_patch = _pair(dilation), _pair(dilation_patch)
    if kernel_size != (1, 1):
        raise NotImplementedError("Only kernel_size=1 is supported.")
    if dilation != (1, 1):
        raise NotImplementedError("Only dilation=1 is supported.")
    if patch_size[0] % 2 == 0 or patch_size[1] % 2 == 0:
        raise NotImplementedError("Only odd patch sizes are supperted.")
    if max(padding) > 0:
        pw = ((0, 0), (0, 0), (padding[0], padding[0]), (padding[1], padding[1]))
        input1 = np.pad(input1, pw)
        input2 = np.pad(input2, pw)
    md = (dilation_patch[0] * (patch_size[0] - 1) // 2, dilation_patch[1] * (patch_size[1] - 1) // 2)
    input2 = np.pad(input2, ((0, 0), (0, 0), (md[0], md[0]), (md[1], md[1])))
    b, _, h, w = input1.shape
    input1 = input1[:, :, :: stride[0], :: stride[1]]
    sh, sw = input1.shape[2:4]
    corr = np.zeros((b, patch_size[0], patch_size[1], sh, sw), dtype=input1.dtype)
    for i in range(0, 2 * md[0] + 1, dilation_patch[0]):
        for j in range(0, 2 * md[1] + 1, dilation_patch[1]):
            p2 = input2[:, :, i : i + h, j : j + w][:, :, :: stride[0], :: stride[1]]
            corr[:, i // dilation_patch[0], j // dilation_patch[1]] = (input1 * p2).sum(axis=1)
    return corr


def iter_spatial_correlation_sample_bwd(input1, input2, dout, patch_size=1, stride=1, padding=0, dilation_patch=1):
    """Adjoint of K4, closed form over the displacement window."""
    patch_size, stride = _pair(patch_size), _pair(stride)
    padding, dp = _pair(padding), _pair(dilation_patch)
    b, c, h0, w0 = input1.shape
    pw = ((0, 0), (0, 0), (padding[0], padding[0]), (padding[1], padding[1]))
    in1 = np.pad(input1, pw).astype(np.float32)
    in2 = np.pad(input2, pw).astype(np.float32)
    md = (dp[0] * (patch_size[0] - 1) // 2, dp[1] * (patch_size[1] - 1) // 2)
    h, w = in1.shape[2:]
    in2p = np.pad(in2, ((0, 0), (0, 0), (md[0], md[0]), (md[1], md[1])))
    d1 = np.zeros_like(in1)
    d2p = np.zeros_like(in2p)
    for pi in range(patch_size[0]):
        for pj in range(patch_size[1]):
            i, j = pi * dp[0], pj * dp[1]
            g = dout[:, pi, pj][:, None].astype(np.float32)  # (b,1,sh,sw)
            p2 = in2p[:, :, i : i + h, j : j + w][:, :, :: stride[0], :: stride[1]]
            d1[:, :, :: stride[0], :: stride[1]] += g * p2
            d2p[:, :, i : i + h, j : j + w][:, :, :: stride[0], :: stride[1]] += g * in1[:, :, :: stride[0], :: stride[1]]
    d2 = d2p[:, :, md[0] : md[0] + h, md[1] : md[1] + w]
    ph, pw_ = padding
    d1 = d1[:, :, ph : ph + h0, pw_ : pw_ + w0]
    d2 = d2[:, :, ph : ph + h0, pw_ : pw_ + w0]
    return d1, d2


# --------------------------------------------------------------------------
# CorrelationLayer                similarity/correlation/layer.py:11-67
# --------------------------------------------------------------------------
def correlation_layer(features1, features2, pad_size: int = 4, max_displacement: int = 4) -> np.ndarray:
    f2p = np.pad(features2, ((0, 0), (0, 0), (pad_size, pad_size), (pad_size, pad_size)))
    h, w = features1.shape[2:]
    k = 2 * max_displacement + 1
    outs = []
    for dy in range(k):  # offsety is the slow index                layer.py:46-64
        for dx in range(k):
            outs.append((features1 * f2p[:, :, dy : dy + h, dx : dx + w]).mean(axis=1, keepdims=True))
    return np.concatenate(outs, axis=1).astype(features1.dtype)


# --------------------------------------------------------------------------
# K6: Matryoshka dilated cost volume      similarity/learnable_cost.py:234-434
# --------------------------------------------------------------------------
def concentric_offsets(dilations: Sequence[int], radius: int) -> np.ndarray:
    """learnable_cost.py:291-298 -> (D, 2r+1) float32."""
    return np.array([(np.arange(-radius, radius + 1) * d).tolist() for d in dilations], dtype=np.float32)


def leaky_relu(x, slope=0.1):
    return np.where(x >= 0, x, x * x.dtype.type(slope)).astype(x.dtype)


def matryoshka_cost_volume(x1, x2, num_groups=1, max_displacement=4, stride=1, dilations=(1, 2, 3, 5, 9, 16), use_relu=False):
    """learnable_cost.py:306-325 -> (B, D*G, 2r+1, 2r+1, h', w')."""
    b, c, h, w = x1.shape
    assert c % num_groups == 0
    cg = c // num_groups
    x1 = x1.reshape(b * num_groups, cg, h, w)
    x2 = x2.reshape(b * num_groups, cg, h, w)
    p = 2 * max_displacement + 1
    costs = []
    for d in dilations:
        cost = iter_spatial_correlation_sample(x1, x2, patch_size=p, stride=stride, padding=0, dilation_patch=d)
        _, u, v, ho, wo = cost.shape
        costs.append(cost.reshape(b, num_groups, u, v, ho, wo))
    cost = np.concatenate(costs, axis=1)
    if use_relu:
        cost = leaky_relu(cost, 0.1)
    return cost


def matryoshka_offsets_2d(max_displacement=4, encoder_output_strides=(2, 8), dilations=((1,), (1, 2, 3, 5, 9, 16))):
    """learnable_cost.py:364-412 -> (sum D, 2r+1, 2r+1, 2) with [...,0]=y, [...,1]=x."""
    rows = []
    for dil, fs in zip(dilations, encoder_output_strides):
        rows.append(concentric_offsets(dil, max_displacement) * np.float32(fs))
    offs = np.concatenate(rows, axis=0)
    nd, sr = offs.shape
    o2 = np.zeros((nd, sr, sr, 2), dtype=np.float32)
    for i in range(nd):
        oi, oj = np.meshgrid(offs[i], offs[i], indexing="ij")
        o2[i, :, :, 0] = oi
        o2[i, :, :, 1] = oj
    return o2


def matryoshka_cost_volume_list(
    x1_list,
    x2_list,
    num_groups=1,
    max_displacement=4,
    encoder_output_strides=(2, 8),
    dilations=((1,), (1, 2, 3, 5, 9, 16)),
    normalize_feat_l2=False,
    use_relu=False,
):
    """learnable_cost.py:420-434."""
    costs = []
    for x1, x2, dil, fs in zip(x1_list, x2_list, dilations, encoder_output_strides):
        if normalize_feat_l2:
            x1 = x1 / (np.sqrt((x1 * x1).sum(axis=1, keepdims=True)) + x1.dtype.type(1e-9))
            x2 = x2 / (np.sqrt((x2 * x2).sum(axis=1, keepdims=True)) + x2.dtype.type(1e-9))
        costs.append(
            matryoshka_cost_volume(x1, x2, num_groups, max_displacement, 8 // fs, dil, use_relu)
        )
    return np.concatenate(costs, axis=1)


# --------------------------------------------------------------------------
# K5: warp                         ezflow/utils/warp.py:5-42
# --------------------------------------------------------------------------
def coords_grid(batch: int, h: int, w: int) -> np.ndarray:
    """ezflow/utils/common.py:9-31 -> (B,2,h,w), ch0 = x index, ch1 = y index."""
    yy, xx = np.meshgrid(np.arange(h), np.arange(w), indexing="ij")
    g = np.stack([xx, yy], axis=0).astype(np.float32)
    return np.repeat(g[None], batch, axis=0)


def warp(x: np.ndarray, flow: np.ndarray) -> np.ndarray:
    b, _, h, w = x.shape
    dt = x.dtype.type
    vgrid = coords_grid(b, h, w).astype(x.dtype) + flow.astype(x.dtype)
    gx = dt(2.0) * vgrid[:, 0] / dt(max(w - 1, 1)) - dt(1.0)
    gy = dt(2.0) * vgrid[:, 1] / dt(max(h - 1, 1)) - dt(1.0)
    grid = np.stack([gx, gy], axis=-1)
    out = grid_sample_bilinear_zeros_ac(x, grid)
    mask = grid_sample_bilinear_zeros_ac(np.ones((b, 1, h, w), dtype=x.dtype), grid)
    mask = np.where(mask < dt(0.9999), dt(0), mask)
    mask = np.where(mask > 0, dt(1), mask)
    return (out * mask).astype(x.dtype)


# --------------------------------------------------------------------------
# K7: convex flow upsampling      ezflow/utils/resampling.py:112-141
# --------------------------------------------------------------------------
def _unfold3x3(flow: np.ndarray) -> np.ndarray:
    """F.unfold(flow, [3,3], padding=1) viewed (N, C, 9, H, W): tap k = ky*3+kx reads flow[h+ky-1, w+kx-1] (zero padded)."""
    n, c, h, w = flow.shape
    padded = np.pad(flow, ((0, 0), (0, 0), (1, 1), (1, 1)))
    return np.stack([padded[:, :, ky : ky + h, kx : kx + w] for ky in range(3) for kx in range(3)], axis=2)


def _softmax(x: np.ndarray, axis: int) -> np.ndarray:
    e = np.exp(x - x.max(axis=axis, keepdims=True))
    return e / e.sum(axis=axis, keepdims=True)


def convex_upsample_flow(flow: np.ndarray, mask_logits: np.ndarray, out_stride: int) -> np.ndarray:
    """resampling.py:131-141: softmax over the 9 taps, weighted 3x3 unfold, pixel-shuffle to (N, C, s*H, s*W)."""
    n, c, h, w = flow.shape
    s = int(out_stride)
    probs = _softmax(mask_logits.reshape(n, 1, 9, s, s, h, w), axis=2)
    taps = _unfold3x3(flow).reshape(n, c, 9, 1, 1, h, w)
    up = (probs * taps).sum(axis=2)  # (n, c, s, s, h, w)
    return up.transpose(0, 1, 4, 2, 5, 3).reshape(n, c, s * h, s * w).astype(flow.dtype)


def convex_upsample_flow_bwd(flow: np.ndarray, mask_logits: np.ndarray, out_stride: int, dout: np.ndarray):
    """Adjoint of convex_upsample_flow (autograd of resampling.py:131-141): returns (dflow, dmask_logits)."""
    n, c, h, w = flow.shape
    s = int(out_stride)
    probs = _softmax(mask_logits.reshape(n, 1, 9, s, s, h, w), axis=2)
    taps = _unfold3x3(flow).reshape(n, c, 9, 1, 1, h, w)
    dup = dout.reshape(n, c, h, s, w, s).transpose(0, 1, 3, 5, 2, 4)[:, :, None]  # (n, c, 1, s, s, h, w)
    g = (dup * taps).sum(axis=1, keepdims=True)  # (n, 1, 9, s, s, h, w)
    dlogits = probs * (g - (probs * g).sum(axis=2, keepdims=True))
    dtaps = (probs * dup).sum(axis=(3, 4))  # (n, c, 9, h, w)
    dpad = np.zeros((n, c, h + 2, w + 2), dtype=flow.dtype)
    for ky in range(3):
        for kx in range(3):
            dpad[:, :, ky : ky + h, kx : kx + w] += dtaps[:, :, ky * 3 + kx]
    return dpad[:, :, 1:-1, 1:-1].astype(flow.dtype), dlogits.reshape(mask_logits.shape).astype(flow.dtype)


# ------------------------------------------------------------------------------------------------
# VCN's per-channel correlation volume and DICL's displacement-stacked volume
# ------------------------------------------------------------------------------------------------
def vcn_corr(features1: np.ndarray, features2: np.ndarray, max_disp: int, factorization: int = 1, slope: float = 0.1) -> np.ndarray:
    """VCN._corr_fn (reference ezflow/models/vcn.py:162-206): cost[b,c,i,j,y,x] = leaky_relu(f1[b,c,y,x] * f2[b,c,y+indd,x+ind]),
    ind = i - max_disp (x shift), indd = j - max_disp // factorization (y shift), zero where the shifted pixel is outside."""
    b, c, h, w = features1.shape
    mv = int(max_disp // factorization)
    cost = np.zeros((b, c, 2 * max_disp + 1, 2 * mv + 1, h, w), dtype=np.float32)
    for i in range(2 * max_disp + 1):
        ind = i - max_disp
        for j in range(2 * mv + 1):
            indd = j - mv
            ys, ye = max(0, -indd), min(h, h - indd)
            xs, xe = max(0, -ind), min(w, w - ind)
            if ys >= ye or xs >= xe:
                continue
            cost[:, :, i, j, ys:ye, xs:xe] = features1[:, :, ys:ye, xs:xe] * features2[:, :, ys + indd : ye + indd, xs + ind : xe + ind]
    return leaky_relu(cost, slope)


def dicl_cost_volume(x: np.ndarray, y: np.ndarray, max_u: int, max_v: int, remove_warp_hole: bool = True) -> np.ndarray:
    """The volume LearnableMatchingCost.forward hands to its matching network (reference
    ezflow/similarity/learnable_cost.py:181-226): (B*U*V, 2C, H, W), image (b,i,j) = [x ; y shifted by (j - max_v, i - max_u)]
    on the overlap, zero elsewhere; with remove_warp_hole also zero where the shifted y sums to 0 over its channels."""
    b, c, h, w = x.shape
    U, V = 2 * max_u + 1, 2 * max_v + 1
    cost = np.zeros((b, 2 * c, U, V, h, w), dtype=np.float32)
    for i in range(U):
        ind = i - max_u
        for j in range(V):
            indd = j - max_v
            ys, ye = max(0, -indd), min(h, h - indd)
            xs, xe = max(0, -ind), min(w, w - ind)
            if ys >= ye or xs >= xe:
                continue
            cost[:, :c, i, j, ys:ye, xs:xe] = x[:, :, ys:ye, xs:xe]
            cost[:, c:, i, j, ys:ye, xs:xe] = y[:, :, ys + indd : ye + indd, xs + ind : xe + ind]
    if remove_warp_hole:
        valid = cost[:, c:].sum(axis=1) != 0
        cost = cost * valid[:, None].astype(np.float32)
    return np.ascontiguousarray(cost.transpose(0, 2, 3, 1, 4, 5)).reshape(b * U * V, 2 * c, h, w)
