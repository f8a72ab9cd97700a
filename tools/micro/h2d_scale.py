#!/usr/bin/env python
"""Copy-only ceiling of bench.py's end-to-end leg: every rank streams the SAME pinned host buffers bench.py uses per step
(8 pairs: fmap1 + fmap2 + 12 coordinate sets in, the last lookup's (8,324,135,240) features out) with NO kernels, H2D and D2H
on two streams so both directions run concurrently, `steps` times.  Prints pairs/s-equivalent and GB/s per direction, max
over ranks — if this curve is as flat from 1 to 8 ranks as the end-to-end number, the limit is the host side of the links.
    python tools/micro/h2d_scale.py [steps]                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/micro/h2d_scale.py [steps]
"""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

steps = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 20
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
try:
    import pynvml

    pynvml.nvmlInit()
    pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(local))
except Exception:
    pass
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=dev)
H, W, C, L, R, ITERS, P = 135, 240, 256, 4, 4, 12, 8
WC = "--wc" in sys.argv  # H2D sources in write-combined pinned memory (cudaHostAllocWriteCombined)


def host_buffer(*shape):
    if not WC:
        return torch.empty(*shape).pin_memory()
    import ctypes

    import numpy as np

    rt = ctypes.CDLL("libcudart.so.12")
    n = 4
    for d in shape:
        n *= d
    p = ctypes.c_void_p()
    rc = rt.cudaHostAlloc(ctypes.byref(p), ctypes.c_size_t(n), ctypes.c_uint(0x04))
    assert rc == 0, f"cudaHostAlloc failed: {rc}"
    arr = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_float)), shape=(n // 4,))
    arr[:] = 0.5
    return torch.from_numpy(arr).view(*shape)


h_in = [host_buffer(P, C, H, W), host_buffer(P, C, H, W)] + [host_buffer(P, 2, H, W) for _ in range(ITERS)]
d_in = [torch.empty_like(t, device=dev) for t in h_in]
d_out = torch.empty(P, L * (2 * R + 1) ** 2, H, W, device=dev)
h_out = torch.empty(d_out.shape).pin_memory()
s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
h2d = sum(t.numel() * 4 for t in h_in)
d2h = d_out.numel() * 4


def run(n, do_in=True, do_out=True):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for _ in range(n):
        if do_in:
            with torch.cuda.stream(s_in):
                for a, b in zip(d_in, h_in):
                    a.copy_(b, non_blocking=True)
        if do_out:
            with torch.cuda.stream(s_out):
                h_out.copy_(d_out, non_blocking=True)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    return time.perf_counter() - t0


run(2)
res = {}
for name, kw in (("both", {}), ("h2d_only", {"do_out": False}), ("d2h_only", {"do_in": False})):
    dt = run(steps, **kw)
    t = torch.tensor([dt], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dt = float(t.item())
    res[name] = {"s": dt, "pairs_per_s_equiv": world * P * steps / dt,
                 "h2d_gbs_per_gpu": h2d * steps / dt / 1e9 if kw.get("do_in", True) else 0.0,
                 "d2h_gbs_per_gpu": d2h * steps / dt / 1e9 if kw.get("do_out", True) else 0.0}
if rank == 0:
    print(json.dumps({"n_gpus": world, "steps": steps, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "cpus": os.cpu_count(),
                      "affinity": len(os.sched_getaffinity(0)), **res}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
