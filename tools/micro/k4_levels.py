"""Per-level timing of the PWC-Net local correlation (config 3 shapes) in both modes: 'tc' (prep + tcgen05) vs 'exact' (fp32 SIMT)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import ezflow_b200 as E
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]
m = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=9, padding=0)
for c, h, w in [(196, 7, 16), (128, 14, 32), (96, 28, 64), (64, 56, 128), (32, 112, 256)]:
    a, b = torch.randn(8, c, h, w, device=dev), torch.randn(8, c, h, w, device=dev)
    r = {}
    for mode in ("tc", "exact"):
        E.set_local_corr_mode(mode)
        r[mode] = timed(lambda: m(a, b))
    nb = 2 * a.numel() * 4 + 8 * 81 * h * w * 4
    print(f"C={c:3d} {h}x{w}: tc {r['tc']*1e3:7.1f} us  exact {r['exact']*1e3:7.1f} us  hbm-bound {nb/6556.8e3:6.1f} us", flush=True)
