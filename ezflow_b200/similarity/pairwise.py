"""RAFT multi-scale all-pairs 4-D correlation on B200 (K1 + K2 + K3).

Drop-in for ``ezflow.similarity.correlation.pairwise.MutliScalePairwise4DCorr`` (reference
ezflow/similarity/correlation/pairwise.py:9-88): same name (sic), constructor arguments,
``from_config`` keys, ``__call__`` signature and output layout.  It is a plain class, not an
``nn.Module``, exactly like the reference.

Deviations, all documented in DESIGN.md §6:
* the pyramid is stored in the library's tiled ("subtile-sequence") layout, by default as fp16 with a
  per-pair power-of-two scale (fp32 selectable with ``set_pyramid_dtype("fp32")``); ``corr_pyramid`` is
  therefore a property that materialises reference-shaped fp32 tensors on demand; at most 4 levels;
* no gradient w.r.t. ``coords`` (RAFT detaches them, reference models/raft.py:115);
* CPU tensors raise ``RuntimeError`` — there is no CPU fallback.
"""

import ctypes
import threading
import os

import torch

from .. import _cabi
from ..config import configurable
from ..registry import SIMILARITY_REGISTRY

_PYRAMID_DTYPE = os.environ.get("EZB_PYRAMID_DTYPE", "fp16").lower()


def set_pyramid_dtype(name):
    """'fp16' (default; halves HBM traffic, ~2e-4 relative storage error) or 'fp32'."""
    global _PYRAMID_DTYPE
    name = str(name).lower()
    if name not in ("fp16", "fp32"):
        raise ValueError("pyramid dtype must be 'fp16' or 'fp32'")
    _PYRAMID_DTYPE = name


def get_pyramid_dtype():
    return _PYRAMID_DTYPE


class PyramidLayout:
    """Geometry of the subtile-sequence pyramid buffer (include/ezflow_b200.h)."""

    def __init__(self, H, W, L, dtype_code):
        h = (ctypes.c_int * L)()
        w = (ctypes.c_int * L)()
        sub_base = (ctypes.c_int * (L + 1))()
        n_tiles, n_groups = ctypes.c_int(), ctypes.c_int()
        gb, pb = ctypes.c_int64(), ctypes.c_int64()
        _cabi.check(
            _cabi.lib().ezb_corr_pyramid_layout(
                H, W, L, dtype_code, h, w, sub_base, ctypes.byref(n_tiles), ctypes.byref(n_groups),
                ctypes.byref(gb), ctypes.byref(pb),
            )
        )
        self.H, self.W, self.L = H, W, L
        self.levels = [(h[i], w[i]) for i in range(L)]
        self.sub_base = [sub_base[i] for i in range(L + 1)]
        self.n_tiles, self.n_groups = n_tiles.value, n_groups.value
        self.group_bytes, self.pair_bytes = gb.value, pb.value
        self.dtype_code = dtype_code
        self.es = 2 if dtype_code == _cabi.EZB_F16 else 4
        self._maps = {}
        self._rec = None

    def record_index(self):
        """(32 lanes, 256 columns) element offsets inside a group block, from the library itself."""
        if self._rec is None:
            fn = _cabi.lib().ezb_corr_pyramid_record_offset
            idx = torch.empty((32, 256), dtype=torch.int64)
            for lane in range(32):
                for n in range(256):
                    off = fn(self.dtype_code, lane, n)
                    assert off >= 0 and off % self.es == 0
                    idx[lane, n] = off // self.es
            self._rec = idx
        return self._rec

    def level_map(self, l):
        """(tile, column) LongTensors of shape (h_l, w_l) for level l, from the library itself."""
        if l not in self._maps:
            h, w = self.levels[l]
            tile = (ctypes.c_int32 * (h * w))()
            col = (ctypes.c_int32 * (h * w))()
            _cabi.check(_cabi.lib().ezb_corr_pyramid_level_map(self.H, self.W, self.L, l, tile, col))
            t = torch.tensor(list(tile), dtype=torch.int64).view(h, w)
            c = torch.tensor(list(col), dtype=torch.int64).view(h, w)
            self._maps[l] = (t, c)
        return self._maps[l]


def grad_layout(H, W, L):
    h = (ctypes.c_int * L)()
    w = (ctypes.c_int * L)()
    pitch = (ctypes.c_int * L)()
    qs = (ctypes.c_int64 * L)()
    _cabi.check(_cabi.lib().ezb_corr_grad_layout(H, W, L, h, w, pitch, qs))
    return [(h[i], w[i], pitch[i], qs[i]) for i in range(L)]


# Gradient pyramids (5.6 GB per 1080p pair) that are all-zero and free for the next backward pass.  With the ranged adjoint
# (ezb_corr_grad_ranges) a pass dirties only the support of its lookup windows and zeroes exactly that again afterwards, so
# the buffer set is handed from step to step instead of being memset (1 ms per pair) every time.  One set is kept, for the
# most recent (device, stream, shape); release_gradient_pool() frees it.
_GRAD_POOL = {}
_GRAD_POOL_LOCK = threading.Lock()


def release_gradient_pool():
    """Frees the cached all-zero gradient pyramid (training with MutliScalePairwise4DCorr keeps one between steps)."""
    with _GRAD_POOL_LOCK:
        _GRAD_POOL.clear()


class _PyramidState:
    """Device buffers of one pyramid plus (lazily) its fp32 gradient accumulator."""

    def __init__(self, fmap1, fmap2, num_levels, dtype_name):
        self.B, self.C, self.H, self.W = fmap1.shape
        self.L = int(num_levels)
        self.dtype_code = _cabi.EZB_F16 if dtype_name == "fp16" else _cabi.EZB_F32
        self.tdtype = torch.float16 if dtype_name == "fp16" else torch.float32
        self.device = fmap1.device
        self.fmap1, self.fmap2 = fmap1, fmap2
        self.N = self.H * self.W
        self.layout = PyramidLayout(self.H, self.W, self.L, self.dtype_code)
        self.pyr = None
        self.deq = None
        self.grad_levels = None
        self.grad_layout = None
        self.grad_task = None  # autograd graph-task id of the backward pass that is filling grad_levels
        self.lookup_coords = []  # coords of every differentiable lookup of this pyramid (forward order)
        self.pending = []  # (dout, coords) of the lookups whose adjoint waits for _BuildFn.backward (ezb_corr_lookup_bwd_multi)
        self.pending_task = None
        self.grad_ranges = None  # int32 support ranges of the gradient pyramid (ezb_corr_grad_ranges), or None: dense

    def build(self, ev_gemm=None):
        """ev_gemm: optional (start, end) raw cudaEvent handles recorded around the GEMM launch (bench only)."""
        lib = _cabi.lib()
        ev0, ev1 = ev_gemm if ev_gemm is not None else (None, None)
        with torch.cuda.device(self.device):
            self.pyr = torch.empty(self.B * self.layout.pair_bytes // self.layout.es, dtype=self.tdtype, device=self.device)
            self.deq = torch.empty(self.B, dtype=torch.float32, device=self.device)
            nbytes = ctypes.c_size_t(0)
            _cabi.check(lib.ezb_corr_pyramid_workspace_bytes(self.B, self.C, self.H, self.W, self.L, ctypes.byref(nbytes)))
            ws = torch.empty(nbytes.value, dtype=torch.uint8, device=self.device)
            _cabi.check(
                lib.ezb_corr_pyramid_fwd_ex(
                    _cabi.ptr(self.fmap1), _cabi.ptr(self.fmap2), self.B, self.C, self.H, self.W, self.L, self.dtype_code,
                    _cabi.ptr(self.pyr), _cabi.ptr(self.deq), _cabi.ptr(ws), nbytes.value, _cabi.stream_ptr(self.device),
                    ev0, ev1,
                )
            )
            # ws is released to the caching allocator here; stream ordering keeps that safe

    def level_tensor(self, l, queries=None):
        """Level l as a reference-shaped (rows, 1, h_l, w_l) tensor in storage dtype (a gathered copy).

        queries: optional 1-D LongTensor of flat row indices b*N+q to extract; default all B*N rows."""
        lay = self.layout
        h, w = lay.levels[l]
        ge = lay.group_bytes // lay.es
        dev = self.device
        tile, col = (t.to(dev).reshape(-1) for t in lay.level_map(l))  # (h*w,)
        rec = lay.record_index().to(dev)  # (32, 256)
        if queries is None:
            queries = torch.arange(self.B * self.N, device=dev)
        queries = queries.to(dev)
        b, q = queries // self.N, queries % self.N
        out = torch.empty((len(queries), h * w), dtype=self.tdtype, device=dev)
        chunk = max(1, (1 << 24) // (h * w))  # bound the index tensors (~128 MB of int64)
        for i0 in range(0, len(queries), chunk):
            bb, qq = b[i0 : i0 + chunk, None], q[i0 : i0 + chunk, None]
            block = (bb * lay.n_tiles + tile[None, :]) * lay.n_groups + qq // 32
            out[i0 : i0 + chunk] = self.pyr[block * ge + rec[(qq % 32).expand(-1, h * w), col[None, :].expand(len(bb), -1)]]
        return out.view(len(queries), 1, h, w)

    def _pool_key(self):
        return (self.device.index, torch.cuda.current_stream(self.device).cuda_stream, self.B, self.H, self.W, self.L)

    def ensure_grad(self, radius):
        # A backward pass that never reached _BuildFn.backward (it raised half-way) leaves its partial sums behind; they must
        # not leak into the next pass.  Every engine run has its own graph-task id.
        task = torch._C._current_graph_task_id() if hasattr(torch._C, "_current_graph_task_id") else None
        if self.grad_levels is not None and task != self.grad_task:
            for g in self.grad_levels:
                g.zero_()
        self.grad_task = task
        if self.grad_levels is None:
            lib = _cabi.lib()
            self.grad_layout = grad_layout(self.H, self.W, self.L)
            rows = self.B * self.N
            ranged = bool(self.lookup_coords) and lib.ezb_corr_grad_ranges_supported(self.C, self.H, self.W, self.L) == 1
            if ranged:
                with _GRAD_POOL_LOCK:
                    self.grad_levels = _GRAD_POOL.pop(self._pool_key(), None)
            if self.grad_levels is None:
                self.grad_levels = [torch.zeros(rows * qs, dtype=torch.float32, device=self.device) for (_, _, _, qs) in self.grad_layout]
            if ranged:
                # the support of this pass: the windows of every lookup recorded in the forward pass (a superset of the
                # lookups whose backward will actually run)
                n = ctypes.c_int64(0)
                _cabi.check(lib.ezb_corr_grad_ranges_count(self.B, self.H, self.W, ctypes.byref(n)))
                self.grad_ranges = torch.empty(n.value, dtype=torch.int32, device=self.device)
                _cabi.check(
                    lib.ezb_corr_grad_ranges(
                        _cabi.ptr_array(self.lookup_coords), len(self.lookup_coords), self.B, self.H, self.W, self.L, int(radius),
                        _cabi.ptr(self.grad_ranges), _cabi.stream_ptr(self.device),
                    )
                )
        return self.grad_levels

    def release_grad(self):
        """After the adjoint GEMMs: zero the support again and hand the (all-zero) buffers to the next pass."""
        if self.grad_ranges is not None and self.grad_levels is not None:
            _cabi.check(
                _cabi.lib().ezb_corr_grad_rezero(
                    _cabi.ptr_array(self.grad_levels), _cabi.ptr(self.grad_ranges), self.B, self.H, self.W, self.L,
                    _cabi.stream_ptr(self.device),
                )
            )
            with _GRAD_POOL_LOCK:
                _GRAD_POOL.clear()  # one set at most
                _GRAD_POOL[self._pool_key()] = self.grad_levels
        self.grad_levels, self.grad_task, self.grad_ranges = None, None, None


class _BuildFn(torch.autograd.Function):
    """fmap1, fmap2 -> token.  The pyramid lives in ``state``; the token only threads autograd."""

    @staticmethod
    def forward(ctx, fmap1, fmap2, state):
        state.build()
        ctx.state = state
        ctx.save_for_backward(fmap1, fmap2)  # autograd's version check catches in-place edits between forward and backward
        return torch.zeros((), dtype=torch.float32, device=fmap1.device)

    @staticmethod
    def backward(ctx, _gtoken):
        st = ctx.state
        fmap1, fmap2 = ctx.saved_tensors
        task = torch._C._current_graph_task_id() if hasattr(torch._C, "_current_graph_task_id") else None
        pending = st.pending if task == st.pending_task else []
        st.pending, st.pending_task = [], None
        if not pending and st.grad_levels is None:  # no lookup contributed a gradient
            return torch.zeros_like(fmap1), torch.zeros_like(fmap2), None
        df1 = torch.empty_like(fmap1)
        df2 = torch.empty_like(fmap2)
        with torch.cuda.device(st.device):
            lib = _cabi.lib()
            if pending:  # K3b for all lookups of the pass (one radius per object)
                glv = st.ensure_grad(pending[0][2])
                _cabi.check(
                    lib.ezb_corr_lookup_bwd_multi(
                        _cabi.ptr_array([p[0] for p in pending]), _cabi.ptr_array([p[1] for p in pending]), len(pending),
                        st.B, st.H, st.W, st.L, pending[0][2], _cabi.ptr_array(glv), _cabi.stream_ptr(st.device),
                    )
                )
                del pending
            nws = ctypes.c_size_t(0)
            _cabi.check(lib.ezb_corr_pyramid_bwd_workspace_bytes(st.B, st.C, st.H, st.W, st.L, ctypes.byref(nws)))
            ws = torch.empty(nws.value, dtype=torch.uint8, device=st.device) if nws.value else None
            _cabi.check(
                lib.ezb_corr_pyramid_bwd_ex(
                    _cabi.ptr_array(st.grad_levels), _cabi.ptr(fmap1), _cabi.ptr(fmap2), st.B, st.C, st.H, st.W, st.L,
                    _cabi.ptr(df1), _cabi.ptr(df2), _cabi.ptr(ws), nws.value, _cabi.ptr(st.grad_ranges), _cabi.stream_ptr(st.device),
                )
            )
            st.release_grad()  # consumed
        return df1, df2, None


def _lookup_forward(state, coords, radius):
    B, _, H, W = coords.shape
    K = 2 * radius + 1
    out = torch.empty((B, state.L * K * K, H, W), dtype=torch.float32, device=coords.device)
    with torch.cuda.device(state.device):
        _cabi.check(
            _cabi.lib().ezb_corr_lookup_fwd(
                _cabi.ptr(state.pyr), _cabi.ptr(state.deq), _cabi.ptr(coords), B, H, W, state.L, radius,
                state.dtype_code, _cabi.ptr(out), _cabi.stream_ptr(state.device),
            )
        )
    return out


class _LookupFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, token, coords, state, radius):
        ctx.state, ctx.radius = state, radius
        ctx.save_for_backward(coords)
        if token is not None and token.requires_grad:
            state.lookup_coords.append(coords.detach())  # the support of the gradient pyramid (ensure_grad)
        return _lookup_forward(state, coords, radius)

    @staticmethod
    def backward(ctx, dout):
        (coords,) = ctx.saved_tensors
        st = ctx.state
        dout = dout.contiguous().float()
        B, _, H, W = coords.shape
        gtoken = dcoords = None
        with torch.cuda.device(st.device):
            if ctx.needs_input_grad[0]:  # towards the feature maps: the lookup adjoint (K3b) into the gradient pyramid
                # Deferred: the build node runs after every lookup of the pass and adds all of them in one launch, so that
                # overlapping windows of a query meet in shared memory instead of in HBM.  A pass that never got that far
                # (it raised) must not leak its entries into the next one: every engine run has its own graph-task id.
                task = torch._C._current_graph_task_id() if hasattr(torch._C, "_current_graph_task_id") else None
                if st.pending and task != st.pending_task:
                    st.pending.clear()
                st.pending_task = task
                st.pending.append((dout, coords, ctx.radius))
                gtoken = torch.zeros((), dtype=torch.float32, device=dout.device)
            if ctx.needs_input_grad[1]:  # towards coords (RAFT detaches them, models/raft.py:115; the reference's op has it)
                dcoords = torch.empty_like(coords)
                _cabi.check(
                    _cabi.lib().ezb_corr_lookup_bwd_coords(
                        _cabi.ptr(st.pyr), _cabi.ptr(st.deq), _cabi.ptr(coords), _cabi.ptr(dout), B, H, W, st.L, ctx.radius,
                        st.dtype_code, _cabi.ptr(dcoords), _cabi.stream_ptr(st.device),
                    )
                )
        return gtoken, dcoords, None, None


@SIMILARITY_REGISTRY.register()
class MutliScalePairwise4DCorr:
    """
    Pairwise 4D correlation at multiple scales (RAFT), computed by the sm_100a kernels.

    Parameters
    ----------
    fmap1, fmap2 : torch.Tensor
        (B, C, H, W) CUDA feature maps, C <= 256
    num_levels : int
        Number of levels in the correlation pyramid
    corr_radius : int
        Radius of the lookup window
    """

    # True in the subclass fuse_model() gives to RAFT: lookups without autograd return a LazyLookup handle
    _lazy_lookup = False

    @configurable
    def __init__(self, fmap1, fmap2, num_levels=4, corr_radius=4):
        self.num_levels = num_levels
        self.corr_radius = corr_radius
        f1 = _cabi.require_cuda_f32(fmap1, "fmap1")
        f2 = _cabi.require_cuda_f32(fmap2, "fmap2")
        if f1.dim() != 4 or f1.shape != f2.shape:
            raise ValueError(f"fmap1/fmap2 must both be (B,C,H,W); got {tuple(fmap1.shape)} and {tuple(fmap2.shape)}")
        if int(num_levels) < 1:
            raise ValueError("num_levels must be >= 1")
        self._empty = f1.shape[0] == 0  # an empty batch yields empty lookups, as the reference's ops do
        if self._empty:
            self._state, self._token, self._shape, self._device = None, None, tuple(f1.shape), f1.device
            return
        self._state = _PyramidState(f1.detach(), f2.detach(), num_levels, _PYRAMID_DTYPE)
        if torch.is_grad_enabled() and (fmap1.requires_grad or fmap2.requires_grad):
            self._token = _BuildFn.apply(f1, f2, self._state)
        else:
            if _GRAD_POOL:  # inference after training: do not sit on the pooled gradient pyramid (5.6 GB per 1080p pair)
                release_gradient_pool()
            self._state.build()
            self._token = None

    @property
    def corr_pyramid(self):
        """Reference-shaped list of (B*N, 1, h_l, w_l) fp32 tensors (pairwise.py:31-41).

        The kernels keep the pyramid in a tiled layout, so these are dequantised *copies* gathered on
        demand (diagnostics / tests); nothing on the lookup path uses them."""
        if self._empty:  # the reference's ops return empty tensors for an empty batch
            _, _, H, W = self._shape
            return [torch.zeros((0, 1, H >> l, W >> l), dtype=torch.float32, device=self._device) for l in range(int(self.num_levels))]
        st = self._state
        scale = st.deq.repeat_interleave(st.N).view(-1, 1, 1, 1)
        return [st.level_tensor(l).float() * scale for l in range(st.L)]

    def __call__(self, coords):
        c = _cabi.require_cuda_f32(coords, "coords")
        if self._empty:
            _, _, H, W = self._shape
            return c.new_zeros((0, self.num_levels * (2 * self.corr_radius + 1) ** 2, H, W))
        st = self._state
        if c.dim() != 4 or c.shape[0] != st.B or c.shape[1] != 2 or c.shape[2] != st.H or c.shape[3] != st.W:
            raise ValueError(f"coords must be ({st.B},2,{st.H},{st.W}); got {tuple(coords.shape)}")
        if torch.is_grad_enabled() and (self._token is not None or c.requires_grad):
            return _LookupFn.apply(self._token, c, st, int(self.corr_radius))
        if self._lazy_lookup:
            from .lookup_conv import LazyLookup

            return LazyLookup(self, c.detach())  # consumed by LookupConv2d (lookup -> convc1 fused), SURVEY §8f-1
        return _lookup_forward(st, c.detach(), int(self.corr_radius))

    def _lookup(self, coords):
        return _lookup_forward(self._state, coords, int(self.corr_radius))

    @classmethod
    def from_config(cls, cfg):
        return {
            "num_levels": cfg.NUM_LEVELS,
            "corr_radius": cfg.CORR_RADIUS,
        }

    @staticmethod
    def corr(fmap1, fmap2):
        """All-pairs correlation volume (B, H, W, 1, H, W) (pairwise.py:78-88), computed on tcgen05: fp16 operands with fp32
        accumulation and fp32 storage, i.e. ~3e-4 relative to the reference's fp32 matmul.  Inference-only:
        the result carries no autograd graph (gradients flow through the class's build + lookup path instead)."""
        f1 = _cabi.require_cuda_f32(fmap1, "fmap1")
        f2 = _cabi.require_cuda_f32(fmap2, "fmap2")
        st = _PyramidState(f1.detach(), f2.detach(), 1, "fp32")
        st.build()
        B, _, H, W = f1.shape
        return st.level_tensor(0).reshape(B, H, W, 1, H, W)


class LazyLookupPairwise4DCorr(MutliScalePairwise4DCorr):
    """The class ``fuse_model`` installs as ``RAFT.similarity_fn``: identical, except that a lookup outside autograd returns
    a ``LazyLookup`` for the update block's ``LookupConv2d`` to evaluate inside its own kernel."""

    _lazy_lookup = True
