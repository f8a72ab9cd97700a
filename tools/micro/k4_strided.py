#!/usr/bin/env python
"""DCVNet's fine cost-volume scale (BASELINE config 4: (4,128,192,384) stride-2 features sampled every 4 pixels, 9x9 window):
strided register-tiled fp32 kernel vs the tensor-core path vs the generic fp32 kernel, and the whole cost-volume list.
Timed as CUDA-graph replays with an L2 flush between them.   python tools/micro/k4_strided.py [out.json]"""
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


rows = []
for (b, c, h, w, s) in [(4, 128, 192, 384, 4), (8, 128, 192, 384, 4), (4, 64, 192, 384, 4), (4, 128, 96, 192, 2)]:
    x1, x2 = torch.randn(b, c, h, w, device=dev), torch.randn(b, c, h, w, device=dev)
    m = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=9, stride=s)
    row = {"shape": [b, c, h, w], "stride": s}
    for name, env, mode in (("strided_rt", "1", "tc"), ("tc", "0", "tc"), ("generic_fp32", "0", "exact")):
        os.environ["EZB_LOCAL_CORR_RT"] = env
        E.set_local_corr_mode(mode)
        row[name + "_ms"] = round(timed_graph(lambda: m(x1, x2)), 4)
    os.environ.pop("EZB_LOCAL_CORR_RT", None)
    E.set_local_corr_mode("tc")
    bytes_ = x2.numel() * 4 + x1.numel() * 4 // (s * s) + b * 81 * (h // s) * (w // s) * 4
    row["algorithmic_mb"] = round(bytes_ / 1e6, 1)
    row["frac_hbm"] = round(bytes_ / (row["strided_rt_ms"] * 1e-3) / 6556.8e9, 3)
    rows.append(row)
    print(json.dumps(row), flush=True)

# the whole DCVNet list (config 4)
x1 = [torch.randn(4, 128, 192, 384, device=dev), torch.randn(4, 256, 48, 96, device=dev)]
x2 = [torch.randn(4, 128, 192, 384, device=dev), torch.randn(4, 256, 48, 96, device=dev)]
ml = E.MatryoshkaDilatedCostVolumeList().to(dev)
row = {"op": "cfg4 MatryoshkaDilatedCostVolumeList"}
for name, env in (("with_strided_rt", None), ("tc_only", "0")):
    if env is None:
        os.environ.pop("EZB_LOCAL_CORR_RT", None)
    else:
        os.environ["EZB_LOCAL_CORR_RT"] = env
    row[name + "_ms"] = round(timed_graph(lambda: ml(x1, x2)), 4)
os.environ.pop("EZB_LOCAL_CORR_RT", None)
rows.append(row)
print(json.dumps(row), flush=True)
if len(sys.argv) > 1:
    json.dump(rows, open(sys.argv[1], "w"), indent=1)
