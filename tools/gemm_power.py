#!/usr/bin/env python
"""Board power / clocks while the pyramid-build GEMM runs back to back for `seconds` (default 5): NVML sampled every 10 ms
from a thread of this process.  Writes a CSV (t_ms, power_w, sm_mhz, mem_mhz, temp_c, reasons_hex) and prints a summary
(JSON) with the average GEMM launch time over the loop — evidence for / against the "energy-bound" reading of the GEMM's
roofline.
    python tools/gemm_power.py [seconds] [csv_path] [mode]
mode: build (default; absmax/convert/pack prep + GEMM, the GEMM is ~92 % of it) | lookup (the lookup kernel alone) | idle
"""
import ctypes
import json
import os
import sys
import threading
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ezflow_b200 as E
from ezflow_b200 import _cabi
from ezflow_b200.similarity import pairwise as PW

seconds = float(sys.argv[1]) if len(sys.argv) > 1 else 5.0
csv_path = sys.argv[2] if len(sys.argv) > 2 else "gpurun_out/gemm_power.csv"
mode = sys.argv[3] if len(sys.argv) > 3 else "build"
dev = torch.device("cuda:0")
H, W, C, L, R, P = 135, 240, 256, 4, 4, 8

import pynvml

pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
limit_w = pynvml.nvmlDeviceGetEnforcedPowerLimit(h) / 1e3
samples, stop = [], threading.Event()


def sampler():
    t0 = time.perf_counter()
    while not stop.is_set():
        try:
            r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
        except Exception:
            r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        samples.append(((time.perf_counter() - t0) * 1e3, pynvml.nvmlDeviceGetPowerUsage(h) / 1e3,
                        pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_MEM),
                        pynvml.nvmlDeviceGetTemperature(h, pynvml.NVML_TEMPERATURE_GPU), int(r)))
        time.sleep(0.010)


g = torch.Generator(device="cpu").manual_seed(1)
f1, f2 = torch.randn(P, C, H, W, generator=g).to(dev), torch.randn(P, C, H, W, generator=g).to(dev)
coords = (E.coords_grid(P, H, W) + torch.randn(P, 2, H, W, generator=g) * 4).clamp_(-16, W + 16).to(dev)
lib = _cabi.lib()
gemm_events = []
orig = PW._PyramidState.build


def build(self, ev_gemm=None):
    a, b = ctypes.c_void_p(), ctypes.c_void_p()
    _cabi.check(lib.ezb_event_create(ctypes.byref(a)))
    _cabi.check(lib.ezb_event_create(ctypes.byref(b)))
    gemm_events.append((a, b))
    return orig(self, ev_gemm=(a, b))


with torch.no_grad():
    fn = E.MutliScalePairwise4DCorr(f1, f2, num_levels=L, corr_radius=R)
    fn(coords)
    torch.cuda.synchronize()
    PW._PyramidState.build = build
    th = threading.Thread(target=sampler, daemon=True)
    th.start()
    time.sleep(0.3)  # idle lead-in
    t_start = time.perf_counter()
    n = 0
    while time.perf_counter() - t_start < seconds:
        if mode == "build":
            fn = E.MutliScalePairwise4DCorr(f1, f2, num_levels=L, corr_radius=R)
        elif mode == "lookup":
            for _ in range(12):
                fn(coords)
        else:
            time.sleep(0.01)
        n += 1
        if n % 4 == 0:
            torch.cuda.synchronize()  # keep the launch queue short so the wall clock tracks the device
    torch.cuda.synchronize()
    t_end = time.perf_counter()
    time.sleep(0.3)
    stop.set()
    th.join()

gemm_ms = []
for a, b in gemm_events:
    v = ctypes.c_float()
    _cabi.check(lib.ezb_event_elapsed_ms(a, b, ctypes.byref(v)))
    gemm_ms.append(v.value)
os.makedirs(os.path.dirname(csv_path) or ".", exist_ok=True)
with open(csv_path, "w") as f:
    f.write("t_ms,power_w,sm_mhz,mem_mhz,temp_c,reasons_hex\n")
    for s in samples:
        f.write("%.1f,%.1f,%d,%d,%d,0x%x\n" % s)
lo, hi = 0.3e3 + 500, 0.3e3 + (t_end - t_start) * 1e3 - 100  # steady part of the loaded window
load = [s for s in samples if lo <= s[0] <= hi]
pw = sorted(s[1] for s in load)
ck = sorted(s[2] for s in load)
mask = 0
for s in load:
    mask |= s[5]
names = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown", 0x80: "hw_power_brake"}
out = {"mode": mode, "seconds": t_end - t_start, "iterations": n, "samples_under_load": len(load), "power_limit_w": limit_w,
       "power_w_median": pw[len(pw) // 2] if pw else None, "power_w_p95": pw[int(len(pw) * 0.95)] if pw else None,
       "power_w_max": pw[-1] if pw else None, "sm_mhz_median": ck[len(ck) // 2] if ck else None, "sm_mhz_min": ck[0] if ck else None,
       "sm_mhz_max_seen": ck[-1] if ck else None, "reasons": sorted(v for k, v in names.items() if mask & k),
       "gemm_launches": len(gemm_ms), "gemm_ms_avg": sum(gemm_ms) / len(gemm_ms) if gemm_ms else None,
       "gemm_ms_first10_avg": sum(gemm_ms[:10]) / max(len(gemm_ms[:10]), 1) if gemm_ms else None,
       "gemm_ms_last10_avg": sum(gemm_ms[-10:]) / max(len(gemm_ms[-10:]), 1) if gemm_ms else None}
print(json.dumps(out))
