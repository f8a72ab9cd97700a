#!/usr/bin/env python
"""PWC-Net levels (BASELINE config 3): tensor-core path vs the register-tiled fp32 kernel vs the generic fp32 kernel, per level,
timed as CUDA-graph replays (the eager numbers are Python-launch-bound).   python tools/micro/k4_rt_levels.py [out.json]"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch

import ezflow_b200 as E

dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timed_graph(fn, iters=20):
    fn(); fn()
    torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        g.replay()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


shapes = [(196, 7, 16), (128, 14, 32), (96, 28, 64), (64, 56, 128), (32, 112, 256), (16, 224, 512), (128, 56, 128), (128, 28, 64), (96, 56, 128)]
rows = []
m = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=9)
mf = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=9)
mf.fused_negative_slope = 0.1
for c, h, w in shapes:
    a, b = torch.randn(8, c, h, w, device=dev), torch.randn(8, c, h, w, device=dev)
    f = torch.randn(8, 2, h, w, device=dev) * 2
    row = {"shape": [8, c, h, w]}
    for name, env, small, mode in (("tc", "0", "0", "tc"), ("rt", "1", "0", "tc"), ("rt_small", "1", "1", "tc"), ("generic_fp32", "0", "0", "exact")):
        os.environ["EZB_LOCAL_CORR_RT"] = env
        os.environ["EZB_LOCAL_CORR_RT_SMALL"] = small
        E.set_local_corr_mode(mode)
        row[name + "_ms"] = round(timed_graph(lambda: m(a, b)), 4)
        if mode == "tc" and small == "0":
            row[name + "_warp_fused_ms"] = round(timed_graph(lambda: mf(a, E.lazy_warp(b, f))), 4)
    E.set_local_corr_mode("tc")
    os.environ.pop("EZB_LOCAL_CORR_RT", None)
    os.environ.pop("EZB_LOCAL_CORR_RT_SMALL", None)
    row["auto_ms"] = round(timed_graph(lambda: m(a, b)), 4)
    rows.append(row)
    print(json.dumps(row), flush=True)
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
