#!/usr/bin/env python
"""Profiling experiments on the GEMM kernel: time it with parts disabled (EZB_GEMM_DEBUG bit mask,
see GemmParams::dbg).  Results with dbg != 0 are WRONG on purpose; this is a diagnosis tool only.
The switches exist only in the hooks build (python -m ezflow_b200.build --hooks, done here before the GPU run: nvcc
cross-compiles without a GPU); the shipped library ignores EZB_GEMM_DEBUG."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from ezflow_b200 import build as _build
os.environ["EZB_LIBRARY"] = _build.build(hooks=True)
import torch
from ezflow_b200 import _cabi
from ezflow_b200.similarity import pairwise as PW

lib = _cabi.lib()
dev = torch.device("cuda:0")
B = int(os.environ.get("PAIRS", "4"))
f1 = torch.randn(B, 256, 135, 240, device=dev)
f2 = torch.randn(B, 256, 135, 240, device=dev)

def ev():
    e = ctypes.c_void_p(); _cabi.check(lib.ezb_event_create(ctypes.byref(e))); return e

for dbg in [int(x) for x in (sys.argv[1:] or ["0"])]:
    os.environ["EZB_GEMM_DEBUG"] = str(dbg)
    ts = []
    for it in range(4):
        st = PW._PyramidState(f1, f2, 4, "fp16")
        a, b = ev(), ev()
        st.build(ev_gemm=(a, b))
        v = ctypes.c_float(); _cabi.check(lib.ezb_event_elapsed_ms(a, b, ctypes.byref(v)))
        ts.append(v.value)
        del st
    print(f"dbg={dbg:3d}  gemm_ms min={min(ts[1:]):.3f} all={['%.3f' % t for t in ts]}", flush=True)
os.environ["EZB_GEMM_DEBUG"] = "0"
