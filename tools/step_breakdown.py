#!/usr/bin/env python
"""Where the headline step's time goes, measured in the running pipeline (not under a profiler): CUDA events between the
phases of `steps` back-to-back steps (prep kernels, GEMM, each of the 12 lookups), 8 pairs of 1080p per step.
    python tools/step_breakdown.py [steps]
"""
import ctypes
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ezflow_b200 as E
from ezflow_b200 import _cabi
from ezflow_b200.similarity import pairwise as PW

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
dev = torch.device("cuda:0")
H, W, C, L, R, ITERS, P = 135, 240, 256, 4, 4, 12, 8
g = torch.Generator(device="cpu").manual_seed(1)
f1, f2 = torch.randn(P, C, H, W, generator=g).to(dev), torch.randn(P, C, H, W, generator=g).to(dev)
grid = E.coords_grid(P, H, W)
coords = [(grid + torch.randn(P, 2, H, W, generator=g) * (2 + k)).clamp_(-16, W + 16).to(dev) for k in range(ITERS)]
lib = _cabi.lib()


def raw_event():
    e = ctypes.c_void_p()
    _cabi.check(lib.ezb_event_create(ctypes.byref(e)))
    return e


gemm_events = []
orig = PW._PyramidState.build


def build(self, ev_gemm=None):
    pair = (raw_event(), raw_event())
    gemm_events.append(pair)
    return orig(self, ev_gemm=pair)


def ev():
    e = torch.cuda.Event(enable_timing=True)
    e.record()
    return e


def step(marks):
    marks.append(ev())
    fn = E.MutliScalePairwise4DCorr(f1, f2, num_levels=L, corr_radius=R)
    marks.append(ev())
    for k in range(ITERS):
        fn(coords[k])
        marks.append(ev())


with torch.no_grad():
    for _ in range(3):
        step([])
    torch.cuda.synchronize()
    # (a) plain loop
    a, b = ev(), None
    for _ in range(steps):
        fn = E.MutliScalePairwise4DCorr(f1, f2, num_levels=L, corr_radius=R)
        for k in range(ITERS):
            fn(coords[k])
    b = ev()
    torch.cuda.synchronize()
    plain = a.elapsed_time(b) / steps
    # (b) instrumented loop
    PW._PyramidState.build = build
    gemm_events.clear()
    all_marks = []
    a = ev()
    for _ in range(steps):
        m = []
        step(m)
        all_marks.append(m)
    b = ev()
    torch.cuda.synchronize()
    instr = a.elapsed_time(b) / steps
    build_ms = sum(m[0].elapsed_time(m[1]) for m in all_marks) / steps
    look_ms = [sum(m[1 + k].elapsed_time(m[2 + k]) for m in all_marks) / steps for k in range(ITERS)]
    gemm_ms = []
    for s, e in gemm_events:
        v = ctypes.c_float()
        _cabi.check(lib.ezb_event_elapsed_ms(s, e, ctypes.byref(v)))
        gemm_ms.append(v.value)
    gemm = sum(gemm_ms) / len(gemm_ms)
    gap_next = sum(all_marks[i][-1].elapsed_time(all_marks[i + 1][0]) for i in range(steps - 1)) / (steps - 1)
print(json.dumps({"ms_per_step_plain": plain, "ms_per_step_instrumented": instr, "build_total_ms": build_ms, "gemm_ms": gemm,
                  "prep_ms": build_ms - gemm, "lookups_ms": look_ms, "lookups_total_ms": sum(look_ms),
                  "between_steps_ms": gap_next}))
