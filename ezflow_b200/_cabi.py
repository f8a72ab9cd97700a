"""ctypes binding of the C-ABI library (include/ezflow_b200.h).

There is deliberately no fallback: if the CUDA library is missing, importing a kernel entry point
raises; if a tensor is not a CUDA tensor, the call raises.  Error codes map to the exception types
the reference raises for the same condition (SURVEY.md §8b "Error convention").
"""

import ctypes
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIBPATH = os.environ.get("EZB_LIBRARY") or os.path.join(_HERE, "_lib", "libezflow_b200.so")  # env override: experiment builds

EZB_F32, EZB_F16 = 0, 1
EZB_MAX_LEVELS = 8

_lock = threading.Lock()
_lib = None

c_int, c_i64, c_f32, c_vp, c_sz = ctypes.c_int, ctypes.c_int64, ctypes.c_float, ctypes.c_void_p, ctypes.c_size_t
_P = ctypes.POINTER

_SIGNATURES = {
    "ezb_version": ([], c_int),
    "ezb_last_error": ([ctypes.c_char_p, c_sz], c_int),
    "ezb_corr_pyramid_layout": (
        [c_int, c_int, c_int, c_int, _P(c_int), _P(c_int), _P(c_int), _P(c_int), _P(c_int), _P(c_i64), _P(c_i64)],
        c_int,
    ),
    "ezb_corr_pyramid_level_map": ([c_int, c_int, c_int, c_int, _P(ctypes.c_int32), _P(ctypes.c_int32)], c_int),
    "ezb_corr_pyramid_record_offset": ([c_int, c_int, c_int], c_i64),
    "ezb_corr_grad_layout": ([c_int, c_int, c_int, _P(c_int), _P(c_int), _P(c_int), _P(c_i64)], c_int),
    "ezb_corr_pyramid_workspace_bytes": ([c_int, c_int, c_int, c_int, c_int, _P(c_sz)], c_int),
    "ezb_corr_pyramid_fwd": ([c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_sz, c_vp], c_int),
    "ezb_corr_pyramid_fwd_ex": (
        [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_sz, c_vp, c_vp, c_vp],
        c_int,
    ),
    "ezb_event_create": ([_P(c_vp)], c_int),
    "ezb_event_destroy": ([c_vp], c_int),
    "ezb_event_elapsed_ms": ([c_vp, c_vp, _P(c_f32)], c_int),
    "ezb_corr_lookup_fwd": ([c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp], c_int),
    "ezb_lookup_conv_packed_weight_bytes": ([c_int, _P(c_sz)], c_int),
    "ezb_lookup_conv_pack_weights": ([c_vp, c_int, c_int, c_int, c_vp, c_vp, c_vp], c_int),
    "ezb_corr_lookup_conv1x1_fwd": ([c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_vp, c_vp], c_int),
    "ezb_corr_lookup_bwd": ([c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, _P(c_vp), c_vp], c_int),
    "ezb_corr_lookup_bwd_multi": ([_P(c_vp), _P(c_vp), c_int, c_int, c_int, c_int, c_int, c_int, _P(c_vp), c_vp], c_int),
    "ezb_corr_lookup_bwd_coords": ([c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp], c_int),
    "ezb_corr_pyramid_bwd_workspace_bytes": ([c_int, c_int, c_int, c_int, c_int, _P(c_sz)], c_int),
    "ezb_corr_pyramid_bwd": ([_P(c_vp), c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_sz, c_vp], c_int),
    "ezb_corr_grad_ranges_supported": ([c_int, c_int, c_int, c_int], c_int),
    "ezb_corr_grad_ranges_count": ([c_int, c_int, c_int, _P(c_i64)], c_int),
    "ezb_corr_grad_ranges": ([_P(c_vp), c_int, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp], c_int),
    "ezb_corr_pyramid_bwd_ex": ([_P(c_vp), c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_sz, c_vp, c_vp], c_int),
    "ezb_corr_grad_rezero": ([_P(c_vp), c_vp, c_int, c_int, c_int, c_int, c_vp], c_int),
    "ezb_local_corr_workspace_bytes": ([c_int, c_int, c_int, c_int, c_int, c_int, _P(c_sz)], c_int),
    "ezb_local_corr_fwd": ([c_vp, c_vp] + [c_int] * 12 + [c_f32, c_f32, c_vp, c_i64, c_vp, c_sz, c_vp], c_int),
    "ezb_warp_local_corr_fwd": ([c_vp, c_vp, c_vp] + [c_int] * 12 + [c_f32, c_f32, c_vp, c_i64, c_vp, c_sz, c_vp], c_int),
    "ezb_local_corr_bwd": ([c_vp, c_vp, c_vp] + [c_int] * 12 + [c_f32, c_i64, c_vp, c_vp, c_vp], c_int),
    "ezb_local_corr_bwd_workspace_bytes": ([c_int, c_int, c_int, c_int, _P(c_sz)], c_int),
    "ezb_local_corr_bwd_ex": ([c_vp, c_vp, c_vp] + [c_int] * 12 + [c_f32, c_i64, c_vp, c_vp, c_vp, c_sz, c_vp], c_int),
    "ezb_matryoshka_bwd": ([c_vp, c_vp, c_vp] + [c_int] * 7 + [_P(c_int), c_int, c_i64, c_vp, c_vp, c_vp, c_sz, c_int, c_vp], c_int),
    "ezb_matryoshka_fwd": (
        [c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, _P(c_int), c_int, c_f32, c_vp, c_i64, c_vp, c_sz, c_vp],
        c_int,
    ),
    "ezb_l2_normalize_fwd": ([c_vp, c_int, c_int, c_int, c_f32, c_vp, c_vp], c_int),
    "ezb_warp_fwd": ([c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp], c_int),
    "ezb_warp_bwd": ([c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp], c_int),
    "ezb_vcn_corr_fwd": ([c_vp, c_vp] + [c_int] * 6 + [c_f32, c_vp, c_vp], c_int),
    "ezb_vcn_corr_bwd": ([c_vp, c_vp, c_vp] + [c_int] * 6 + [c_f32, c_vp, c_vp, c_vp], c_int),
    "ezb_dicl_volume_fwd": ([c_vp, c_vp] + [c_int] * 7 + [c_vp, c_vp, c_vp], c_int),
    "ezb_dicl_volume_bwd": ([c_vp, c_vp] + [c_int] * 6 + [c_vp, c_vp, c_vp], c_int),
    "ezb_corr_ondemand_workspace_bytes": ([c_int] * 5 + [_P(c_sz)], c_int),
    "ezb_corr_ondemand_prepare": ([c_vp, c_vp] + [c_int] * 5 + [c_vp, c_sz, c_vp], c_int),
    "ezb_corr_ondemand_lookup": ([c_vp, c_vp] + [c_int] * 6 + [c_vp, c_vp], c_int),
    "ezb_convex_upsample_fwd": ([c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp], c_int),
    "ezb_convex_upsample_bwd_workspace_bytes": ([c_int, c_int, c_int, c_int, _P(c_sz)], c_int),
    "ezb_convex_upsample_bwd": ([c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_vp, c_vp, c_vp, c_sz, c_vp], c_int),
}
EXPORTED_SYMBOLS = tuple(_SIGNATURES)


def lib():
    """Loads the shared library once.  Fails loudly when it has not been built."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIBPATH):
                    raise RuntimeError(
                        f"ezflow_b200: CUDA library not found at {LIBPATH}. Build it with "
                        "`python -m ezflow_b200.build` (nvcc, sm_100a). There is no CPU fallback."
                    )
                l = ctypes.CDLL(LIBPATH)
                for name, (argtypes, restype) in _SIGNATURES.items():
                    fn = getattr(l, name)
                    fn.argtypes = argtypes
                    fn.restype = restype
                _lib = l
    return _lib


def last_error():
    buf = ctypes.create_string_buffer(512)
    lib().ezb_last_error(buf, 512)
    return buf.value.decode(errors="replace")


def check(rc):
    if rc == 0:
        return
    msg = last_error()
    if rc == -1:
        raise ValueError(f"ezflow_b200: {msg}")
    if rc == -2:
        raise NotImplementedError(msg)
    raise RuntimeError(f"ezflow_b200: {msg} (status {rc})")


def require_cuda_f32(t, name):
    if not isinstance(t, torch.Tensor):
        raise TypeError(f"{name} must be a torch.Tensor, got {type(t).__name__}")
    if not t.is_cuda:
        raise RuntimeError(
            f"ezflow_b200: {name} is on {t.device}; the B200 similarity path has no CPU fallback (move inputs to a CUDA device)"
        )
    if t.dtype != torch.float32:
        t = t.float()  # the reference computes this path in fp32 (models/raft.py:93-94)
    return t.contiguous()


def stream_ptr(device):
    return c_vp(torch.cuda.current_stream(device).cuda_stream)


def ptr(t):
    return c_vp(t.data_ptr()) if t is not None else c_vp(0)


def ptr_array(tensors):
    arr = (c_vp * len(tensors))()
    for i, t in enumerate(tensors):
        arr[i] = t.data_ptr()
    return arr
