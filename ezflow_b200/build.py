"""Builds the C-ABI shared library in-tree:  ezflow_b200/_lib/libezflow_b200.so

    python -m ezflow_b200.build [--force]

nvcc cross-compiles sm_100a without a GPU.  The .so is git-ignored but travels to the GPU box.
"""

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "_lib")
LIBPATH = os.path.join(LIBDIR, "libezflow_b200.so")
SOURCES = ["cabi.cu", "corr_pyramid.cu", "corr_lookup.cu", "corr_lookup_conv.cu", "corr_bwd.cu", "corr_bwd_tc.cu", "local_corr.cu", "local_corr_rt.cu", "local_corr_tc.cu", "local_corr_bwd_tc.cu", "warp.cu",
           "convex_upsample.cu", "volume_ops.cu"]
NVCC_FLAGS = [
    "-std=c++17", "-O3", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC",
]


def _nvcc():
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    raise RuntimeError("nvcc not found")


def _digest():
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for root in (CSRC, os.path.join(os.path.dirname(HERE), "include")):
        for f in sorted(os.listdir(root)):
            if f.endswith((".cu", ".cuh", ".h")):
                with open(os.path.join(root, f), "rb") as fh:
                    h.update(f.encode())
                    h.update(fh.read())
    return h.hexdigest()


def build(force=False, verbose=False, hooks=False):
    """hooks=True builds libezflow_b200_hooks.so with -DEZB_PROFILING_HOOKS (the GEMM's EZB_GEMM_DEBUG experiment switches,
    tools/gemm_dbg.py); the shipped library is built without them."""
    os.makedirs(LIBDIR, exist_ok=True)
    libpath = LIBPATH.replace(".so", "_hooks.so") if hooks else LIBPATH
    stamp = os.path.join(LIBDIR, "build_hooks.sha256" if hooks else "build.sha256")
    dig = _digest()
    if not force and os.path.exists(libpath) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return libpath
    nvcc = _nvcc()
    objs = []
    procs = []
    extra = ["-DEZB_PROFILING_HOOKS"] if hooks else []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace(".cu", "_hooks.o" if hooks else ".o"))
        cmd = [nvcc] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-c", os.path.join(CSRC, src), "-o", obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    for src, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode:
            sys.stderr.write(out)
        if p.returncode:
            raise RuntimeError(f"nvcc failed on {src}")
    cmd = [nvcc, "-shared", "-o", libpath] + objs + ["-gencode", "arch=compute_100a,code=sm_100a"]
    subprocess.check_call(cmd)
    with open(stamp, "w") as f:
        f.write(dig)
    return libpath


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv, hooks="--hooks" in sys.argv))
