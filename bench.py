#!/usr/bin/env python
"""Headline benchmark: RAFT 1080p correlation-pyramid build + 12 lookups, image pairs / second.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N ...            # the unmodified reference's own class on the host cores

One "step" = every rank builds the 4-level all-pairs correlation pyramid (K1+K2) of `--pairs-per-gpu`
synthetic 1080x1920 pairs (feature maps 256x135x240 fp32, random normal) and runs the 12 radius-4
lookups (K3) RAFT performs per forward (reference ezflow/models/raft.py:96-116).  Pairs are sharded over
ranks along the batch dimension with no data-path collective (weak scaling: fixed work per GPU).
Prints ONE JSON line on rank 0 (contract in the task statement; fields documented in DESIGN.md §7).
"""

import argparse
import ctypes
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "image pairs/sec (RAFT 1080p corr+12 lookups)"
H, W, C, L, R, ITERS = 135, 240, 256, 4, 4, 12
N = H * W
LEVEL_ELEMS = 32400 + 8040 + 1980 + 480  # 135x240 + 67x120 + 33x60 + 16x30 per query


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs-per-gpu", type=int, default=8, help="image pairs each rank processes per step")
    ap.add_argument("--micro", type=int, default=8, help="pairs per pyramid build (micro-batch)")
    ap.add_argument("--pyramid-dtype", default="fp16", choices=["fp16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-sample-rows", type=int, default=4, help="query rows (of 135) per worker in the CPU-port sample")
    ap.add_argument("--no-secondary", action="store_true", help="skip the BASELINE config 2-5 records")
    ap.add_argument("--secondary", default="", help="comma-separated subset of secondary records to run (default: all)")
    ap.add_argument("--train-batch", type=int, default=64, help="global batch of the config-5 training step (sharded over ranks)")
    ap.add_argument("--train-micro", type=int, default=4, help="pairs per pyramid build in the config-5 training step")
    ap.add_argument("--sustained-seconds", type=float, default=2.0, help="length of the sustained-clock headline record")
    return ap.parse_args()


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d["hbm_gbs"], "tc_tflops": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "src": "measured"}
    return {"hbm_gbs": 6650.0, "tc_tflops": 1400.0, "src": "fallback"}


# ----------------------------------------------------------------------------------------------
# CPU port of the reference path (oracle) on a bounded sample: a band of query rows of one pair.
# Queries are independent in this workload, so pairs/s = (sample queries / N) / seconds.
# ----------------------------------------------------------------------------------------------
def _cpu_worker(args):
    seed, rows = args
    import numpy as np
    from threadpoolctl import threadpool_limits

    from oracle import similarity_oracle as O

    rng = np.random.default_rng(seed)
    f1s = rng.standard_normal((1, C, rows, W)).astype(np.float32)  # the sampled band of fmap1
    f2 = rng.standard_normal((1, C, H, W)).astype(np.float32)
    coords = [
        np.clip(O.coords_grid(1, rows, W) + rng.standard_normal((1, 2, rows, W)).astype(np.float32) * (2 + k), -16, W + 16)
        for k in range(ITERS)
    ]
    with threadpool_limits(limits=1):
        t0 = time.perf_counter()
        nq = rows * W
        corr = np.matmul(f1s.reshape(C, nq).T, f2.reshape(C, N)) / np.float32(np.sqrt(np.float32(C)))
        pyr = [corr.reshape(nq, 1, H, W)]
        for _ in range(L - 1):
            pyr.append(O.avg_pool2x2(pyr[-1]))
        acc = 0.0
        for k in range(ITERS):
            acc += float(O.corr_lookup(pyr, coords[k], R).sum())
        dt = time.perf_counter() - t0
    return dt, acc


def cpu_port_pairs_per_sec(rows_per_worker, workers=None, repeats=1):
    """Times the numpy restatement with `workers` processes (one per host core), each handling its own
    band of `rows_per_worker` query rows.  Returns (pairs/s, cores, sample description)."""
    import multiprocessing as mp

    try:  # the GPU leg may have pinned this process next to its GPU: the CPU baseline gets every host core
        os.sched_setaffinity(0, range(os.cpu_count() or 1))
    except (AttributeError, OSError):
        pass
    workers = workers or os.cpu_count() or 1
    ctx = mp.get_context("fork")
    best = None
    with ctx.Pool(workers) as pool:
        for rep in range(repeats):
            res = pool.map(_cpu_worker, [(1000 + rep * workers + i, rows_per_worker) for i in range(workers)])
            wall = max(r[0] for r in res)  # workers run concurrently; the slowest one bounds the step
            best = wall if best is None else min(best, wall)
    queries = workers * rows_per_worker * W
    value = (queries / N) / best
    sample = (
        f"numpy port of pairwise.py on {workers} processes x {rows_per_worker} query rows of one 1080p pair "
        f"({queries} of {N} queries: fp32 GEMM vs full fmap2, 3 avg-pools, 12 lookups), wall {best:.2f}s, "
        f"extrapolated by query count"
    )
    return value, workers, sample


def _import_reference_class():
    """The UNMODIFIED reference's MutliScalePairwise4DCorr (pip-installed under git-ignored baseline/_ref by
    __graft_entry__.build(); /root/reference where it exists), through the fvcore/omegaconf shim.  None when absent."""
    try:
        tests_dir = os.path.join(ROOT, "tests")
        if tests_dir not in sys.path:
            sys.path.insert(0, tests_dir)
        import warnings

        import refshim

        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            if refshim.import_reference() is None:
                return None
            from ezflow.similarity import SIMILARITY_REGISTRY

            cls = SIMILARITY_REGISTRY.get("MutliScalePairwise4DCorr")
        return cls if cls.__module__.startswith("ezflow.") else None
    except Exception as e:  # noqa: BLE001 — the caller falls back to the numpy port and says so
        sys.stderr.write(f"bench: reference import failed ({e!r}); using the numpy port\n")
        return None


class ReferenceCpuArm:
    """One step = the reference's own class on ONE full 1080p pair on the host: build (matmul + 3 avg_pool2d) + 12 lookups
    (grid_sample), torch CPU ops on every host core — the bounded sample of the workload (about 3-5 s, 9 GB RSS)."""

    def __init__(self):
        import torch

        self.torch = torch
        self.cls = _import_reference_class()
        try:
            os.sched_setaffinity(0, range(os.cpu_count() or 1))
        except (AttributeError, OSError):
            pass
        self.cores = os.cpu_count() or 1
        torch.set_num_threads(self.cores)
        if self.cls is not None:
            g = torch.Generator().manual_seed(4321)
            self.f1 = torch.randn(1, C, H, W, generator=g)
            self.f2 = torch.randn(1, C, H, W, generator=g)
            ys, xs = torch.meshgrid(torch.arange(H, dtype=torch.float32), torch.arange(W, dtype=torch.float32), indexing="ij")
            grid = torch.stack([xs, ys], 0)[None]
            self.coords = [(grid + torch.randn(1, 2, H, W, generator=g) * (2 + k)).clamp_(-16, W + 16) for k in range(ITERS)]

    @property
    def kind(self):
        return "reference" if self.cls is not None else "port"

    def step(self):
        """Seconds for one pair."""
        torch = self.torch
        t0 = time.perf_counter()
        with torch.no_grad():
            fn = self.cls(self.f1, self.f2, num_levels=L, corr_radius=R)
            acc = 0.0
            for k in range(ITERS):
                acc += float(fn(self.coords[k])[0, 0, 0, 0])
        return time.perf_counter() - t0

    def sample(self):
        return (f"ezflow.similarity.MutliScalePairwise4DCorr (unmodified reference, torch {self.torch.__version__} CPU ops, "
                f"{self.torch.get_num_threads()} threads) on ONE full 1080p pair per step: build + 12 lookups, fp32")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    arm = ReferenceCpuArm()
    times = []
    if arm.cls is not None:
        for s in range(args.warmup + args.steps):
            dt = arm.step()
            if s >= args.warmup:
                times.append(dt)
        total = sum(times)
        value = len(times) / total
        workers, sample = arm.cores, "each step: " + arm.sample()
    else:  # no reference on this machine: the numpy port on a band of query rows
        workers = os.cpu_count() or 1
        import multiprocessing as mp

        ctx = mp.get_context("fork")
        with ctx.Pool(workers) as pool:
            for s in range(args.warmup + args.steps):
                res = pool.map(_cpu_worker, [(s * workers + i, args.cpu_sample_rows) for i in range(workers)])
                dt = max(r[0] for r in res)  # workers run concurrently; the slowest one bounds the step
                if s >= args.warmup:
                    times.append(dt)
        queries = workers * args.cpu_sample_rows * W
        total = sum(times)
        value = (queries / N) * len(times) / total
        sample = (
            f"each step: numpy port of pairwise.py on {workers} processes x {args.cpu_sample_rows} query rows of one 1080p pair "
            f"({queries} of {N} queries; fp32 GEMM vs full fmap2, 3 avg-pools, 12 lookups), extrapolated by query count"
        )
    line = {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": "pairs/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": "raft_1080p_corr_build_12_lookups", "feat": [C, H, W], "levels": L, "radius": R, "iters": ITERS,
                   "pairs_per_step": 1 if arm.cls is not None else None},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": workers, "kind": arm.kind, "sample": sample},
        "e2e": {"value": value, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ----------------------------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled every 5 ms DURING the timed region through NVML (a thread in this
    process; falls back to one nvidia-smi query when NVML is unavailable)."""

    REASONS = {
        "hw_slowdown": 0x8, "sw_power_cap": 0x4, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20,
        "hw_power_brake_slowdown": 0x80,
    }

    def __init__(self, index):
        import threading

        self.index = index
        self.samples, self.mask, self.max_mhz = [], 0, None
        self._stop = threading.Event()
        self._nvml = None
        try:
            import pynvml

            pynvml.nvmlInit()
            # NVML enumerates physical devices: map the CUDA ordinal through CUDA_VISIBLE_DEVICES when it is set
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
        except Exception:
            self._nvml = None
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()

    def _run(self):
        nv = self._nvml
        if nv is None:
            return
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                try:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                except Exception:
                    self.mask |= int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self._h))
            except Exception:
                break
            time.sleep(0.005)

    def stop(self):
        self._stop.set()
        self._t.join(timeout=2)
        out = {"sm_mhz": None, "sm_max_mhz": self.max_mhz, "reasons": [], "samples": len(self.samples), "source": "nvml"}
        if self.samples:
            s = sorted(self.samples)
            out["sm_mhz"] = s[len(s) // 2]
            out["sm_min_mhz"] = s[0]
            out["reasons"] = sorted(k for k, bit in self.REASONS.items() if self.mask & bit)
            return out
        # fallback: a single nvidia-smi query (after the region; reported as such)
        out["source"] = "nvidia-smi (single query after the timed region)"
        try:
            q = "clocks.sm,clocks.max.sm"
            r = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                               capture_output=True, text=True, timeout=10).stdout.strip().split(",")
            out["sm_mhz"], out["sm_max_mhz"], out["samples"] = float(r[0]), float(r[1]), 1
        except Exception:
            pass
        return out


def _bind_to_gpu_numa_node(index):
    """Pins this process to the CPUs NVML reports as closest to GPU `index` (no-op when NVML cannot tell)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        phys = index
        if vis:
            ids = [v.strip() for v in vis.split(",") if v.strip()]
            if index < len(ids) and ids[index].isdigit():
                phys = int(ids[index])
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(phys))
    except Exception:
        pass


# ----------------------------------------------------------------------------------------------
# Secondary records: BASELINE configs 2-5 (+ a sustained-clock run of the headline step), driver-observed.
# Every record: median (or total) device time from CUDA events, a 256 MB buffer zeroed between timed iterations
# (L2 flush), algorithmic bytes from BASELINE.md §3 / DESIGN.md §5, fraction of the measured HBM peak, and the
# DRAM traffic of the same launches from the committed ncu capture (profiles/secondary_traffic.json) when there is one.
# Under torchrun every rank runs its own copy (batch-sharded, weak) and the time is the max over ranks; config 5 is the
# one strong-scaling record: a global batch of 64 pairs sharded over the ranks + one NCCL all-reduce of a RAFT-sized
# gradient bucket per step (reference ezflow/engine/trainer.py:269-299: DDP all-reduce inside backward()).
# ----------------------------------------------------------------------------------------------
RAFT_PARAMS = 5_257_536  # parameters of configs/models/raft.yaml = elements of the fp32 gradient bucket DDP reduces


def run_secondary(args, dev, world, rank, peaks, headline_step):
    import torch

    import ezflow_b200 as E
    from ezflow_b200.sharding import max_over_ranks, shard_range

    want = [w for w in args.secondary.split(",") if w]
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "secondary_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath))
        except Exception:
            traffic = {}
    records = []

    def timed(fn, iters=10, warm=3):
        for _ in range(warm):
            fn()
        ts = []
        for _ in range(iters):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ts.sort()
        (ms,) = max_over_ranks([ts[len(ts) // 2]], device=dev)
        return ms

    def timed_graph(fn, iters=10):
        """The same call sequence captured once into a CUDA graph and replayed: device time without the Python / ctypes launch
        path, which dominates eager timings of ops that take tens of microseconds.  None when capture is not possible."""
        try:
            side = torch.cuda.Stream(dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                fn()
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                fn()
            for _ in range(3):
                g.replay()
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
            (ms,) = max_over_ranks([ts[len(ts) // 2]], device=dev)
            del g
            return ms
        except Exception as e:  # noqa: BLE001
            sys.stderr.write(f"bench: CUDA-graph capture failed ({e!r})\n")
            torch.cuda.synchronize()
            return None

    def add(name, ms, nbytes, note, **extra):
        rec = {"name": name, "ms": ms, "algorithmic_bytes": nbytes, "gbs": nbytes / ms / 1e6 if ms > 0 else None,
               "frac": (nbytes / ms / 1e6) / peaks["hbm_gbs"] if ms > 0 else None, "bound": "hbm", "traffic": traffic.get(name),
               "note": note}
        rec.update(extra)
        if rec.get("ms_graph"):
            rec["frac_graph"] = (nbytes / rec["ms_graph"] / 1e6) / peaks["hbm_gbs"]
        records.append(rec)

    OPT_IN = {"cfg4_dcvnet_matryoshka_bwd"}  # not a BASELINE configuration: only when named in --secondary

    def enabled(name):
        return (name in want) if (want or name in OPT_IN) else True

    with torch.no_grad():
        if enabled("cfg2_flownetc_corr_fwd") or enabled("cfg2_flownetc_corr_bwd"):
            x1, x2 = torch.randn(8, 256, 48, 64, device=dev), torch.randn(8, 256, 48, 64, device=dev)
            m = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=21, padding=0, dilation_patch=2)
            nb_f = 2 * x1.numel() * 4 + 8 * 441 * 48 * 64 * 4
            t_f = timed(lambda: m(x1, x2))
            if enabled("cfg2_flownetc_corr_fwd"):
                add("cfg2_flownetc_corr_fwd", t_f, nb_f, "IterSpatialCorrelationSampler P=21 dilation_patch=2 on (8,256,48,64) x2 (BASELINE config 2)",
                    ms_graph=timed_graph(lambda: m(x1, x2)))
            if enabled("cfg2_flownetc_corr_bwd"):
                with torch.enable_grad():
                    a, b = x1.clone().requires_grad_(True), x2.clone().requires_grad_(True)
                    dout = torch.randn(8, 21, 21, 48, 64, device=dev)

                    def fb():
                        a.grad = b.grad = None
                        m(a, b).backward(dout)

                    t_fb = timed(fb)
                add("cfg2_flownetc_corr_bwd", t_fb - t_f, 2 * 2 * x1.numel() * 4 + dout.numel() * 4,
                    "both adjoints of config 2 (forward+backward through autograd minus the forward alone)")
            del x1, x2
        if enabled("cfg3_pwc_corr_fwd") or enabled("cfg3_pwc_warp_corr_leaky_fused") or enabled("cfg3_pwc_finest_bwd"):
            shapes = [(196, 7, 16), (128, 14, 32), (96, 28, 64), (64, 56, 128), (32, 112, 256)]
            xs = [(torch.randn(8, c, h, w, device=dev), torch.randn(8, c, h, w, device=dev), torch.randn(8, 2, h, w, device=dev) * 2) for c, h, w in shapes]
            m = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=9, padding=0)
            nb = sum(2 * a.numel() * 4 + 8 * 81 * a.shape[2] * a.shape[3] * 4 for a, _, _ in xs)
            if enabled("cfg3_pwc_corr_fwd"):
                add("cfg3_pwc_corr_fwd", timed(lambda: [m(a, b) for a, b, _ in xs]), nb,
                    "IterSpatialCorrelationSampler P=9 on the 5 PWC-Net levels, batch 8, 448x1024 input (BASELINE config 3, no warp)",
                    ms_graph=timed_graph(lambda: [m(a, b) for a, b, _ in xs]))
            if enabled("cfg3_pwc_warp_corr_leaky_fused"):
                mf = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=9, padding=0)
                mf.fused_negative_slope = 0.1
                add("cfg3_pwc_warp_corr_leaky_fused", timed(lambda: [mf(a, E.lazy_warp(b, f) if i > 0 else b) for i, (a, b, f) in enumerate(xs)]),
                    nb + sum(f.numel() * 4 for _, _, f in xs[1:]),
                    "PWC decoder step as one kernel chain per level: warp -> correlation -> LeakyReLU(0.1) (reference decoder/pyramid.py:98-102,138-141)",
                    ms_graph=timed_graph(lambda: [mf(a, E.lazy_warp(b, f) if i > 0 else b) for i, (a, b, f) in enumerate(xs)]))
            if enabled("cfg3_pwc_finest_bwd"):
                a0, b0, _ = xs[-1]
                t_f = timed(lambda: m(a0, b0))
                with torch.enable_grad():
                    a, b = a0.clone().requires_grad_(True), b0.clone().requires_grad_(True)
                    dout = torch.randn(8, 9, 9, 112, 256, device=dev)

                    def fb3():
                        a.grad = b.grad = None
                        m(a, b).backward(dout)

                    t_fb = timed(fb3)
                add("cfg3_pwc_finest_bwd", t_fb - t_f, 2 * 2 * a0.numel() * 4 + dout.numel() * 4, "both adjoints of the finest PWC level (8,32,112,256), P=9")
            del xs
        if enabled("cfg4_dcvnet_matryoshka_fwd") or enabled("cfg4_dcvnet_matryoshka_bwd"):
            xa = [torch.randn(4, 128, 192, 384, device=dev), torch.randn(4, 256, 48, 96, device=dev)]
            xb = [torch.randn(4, 128, 192, 384, device=dev), torch.randn(4, 256, 48, 96, device=dev)]
            mm = E.MatryoshkaDilatedCostVolumeList().to(dev)
            add("cfg4_dcvnet_matryoshka_fwd", timed(lambda: mm(xa, xb)), int(240.0e6),
                "MatryoshkaDilatedCostVolumeList on (4,128,192,384)+(4,256,48,96) feature pairs (BASELINE config 4)",
                ms_graph=timed_graph(lambda: mm(xa, xb)))
            if enabled("cfg4_dcvnet_matryoshka_bwd"):
                t_f = timed(lambda: mm(xa, xb))
                with torch.enable_grad():
                    ga = [t.clone().requires_grad_(True) for t in xa]
                    gb = [t.clone().requires_grad_(True) for t in xb]
                    dout = torch.randn(4, 7, 9, 9, 48, 96, device=dev)

                    def fb4():
                        for t in ga + gb:
                            t.grad = None
                        mm(ga, gb).backward(dout)

                    t_fb = timed(fb4)
                add("cfg4_dcvnet_matryoshka_bwd", t_fb - t_f, 2 * 2 * sum(t.numel() for t in xa) * 4 + dout.numel() * 4,
                    "both adjoints of config 4 (ezb_matryoshka_bwd per scale: tensor-core adjoints for the dilations that fit, fp32 kernels for the rest)")
            del xa, xb
    torch.cuda.empty_cache()

    if enabled("lookup_convc1_fused") or enabled("lookup_then_convc1_unfused"):
        # one RAFT iteration's lookup + MotionEncoder.convc1 (1x1, 324 -> 256) + ReLU for 8 1080p pairs
        with torch.no_grad():
            Pn = args.micro
            g = torch.Generator(device="cpu").manual_seed(77 + rank)
            f1 = torch.randn(Pn, C, H, W, generator=g).to(dev)
            f2 = torch.randn(Pn, C, H, W, generator=g).to(dev)
            coords = (E.coords_grid(Pn, H, W) + torch.randn(Pn, 2, H, W, generator=g) * 4).clamp_(-16, W + 16).to(dev)
            wgt = (torch.randn(256, L * (2 * R + 1) ** 2, 1, 1, generator=g) * 0.05).to(dev)
            bias = torch.randn(256, generator=g).to(dev)
            fn = E.LazyLookupPairwise4DCorr(f1, f2, num_levels=L, corr_radius=R)
            lazy = fn(coords)
            E.lookup_conv1x1(lazy, wgt, bias)  # packs the weights once
            # algorithmic bytes: windows (2r+2)^2 fp16 per level + coords in, 256 fp32 channels out
            nb = Pn * N * (L * (2 * R + 2) ** 2 * 2 + 8 + 256 * 4)
            if enabled("lookup_convc1_fused"):
                add("lookup_convc1_fused", timed(lambda: E.lookup_conv1x1(lazy, wgt, bias)), nb,
                    "lookup -> convc1 (1x1, 324->256) + bias + ReLU as ONE tcgen05 kernel, %d 1080p pairs; the (B,324,H,W) features never reach HBM" % Pn)
            if enabled("lookup_then_convc1_unfused"):
                nb_un = nb + 2 * Pn * N * L * (2 * R + 1) ** 2 * 4  # + write and re-read of the lookup features
                add("lookup_then_convc1_unfused", timed(lambda: torch.relu_(torch.nn.functional.conv2d(fn._lookup(coords), wgt, bias))), nb_un,
                    "same result from this repo's lookup kernel + torch conv2d (cuDNN/cuBLAS, TF32 off) + in-place ReLU")
            del fn, lazy, f1, f2
        torch.cuda.empty_cache()

    if enabled("ondemand_lookup_fp32"):
        # the volume-free lookup (no 2.9 GB pyramid per pair): prepare (pooled fmap2 levels) + one lookup of 2 pairs
        with torch.no_grad():
            g = torch.Generator(device="cpu").manual_seed(5 + rank)
            f1 = torch.randn(2, C, H, W, generator=g).to(dev)
            f2 = torch.randn(2, C, H, W, generator=g).to(dev)
            flow = torch.nn.functional.interpolate(torch.randn(2, 2, 5, 8, generator=g) * 10, size=(H, W), mode="bicubic", align_corners=True)
            coords = (E.coords_grid(2, H, W) + flow + torch.randn(2, 2, H, W, generator=g) * 0.5).to(dev)  # a smooth flow + 0.5 px noise
            od = E.OnDemandPairwise4DCorr(f1, f2, num_levels=L, corr_radius=R)
            nb = 2 * N * (L * (2 * R + 1) ** 2 * 4 + 8) + 2 * 2 * C * N * 4  # lookup features out + coords + both feature maps in
            add("ondemand_lookup_fp32", timed(lambda: od(coords)), nb,
                "OnDemandPairwise4DCorr: one lookup of 2 1080p pairs straight from the feature maps (fp32 CUDA cores, 3.3 GFMA per pair, "
                "shared-memory tiled: neighbouring queries share the fmap2 pixels of their windows; coords = grid + smooth flow + 0.5 px noise; "
                "the path for inputs whose all-pairs volume does not fit)", ms_per_pair_lookup=None)
            records[-1]["ms_per_pair_lookup"] = records[-1]["ms"] / 2
            del od, f1, f2, coords
        torch.cuda.empty_cache()

    for c5_name in ("cfg5_raft_1080p_train_step", "cfg5_raft_1080p_train_step_smooth_flow"):
        if not enabled(c5_name):
            continue
        smooth = c5_name.endswith("smooth_flow")
        # per pair: build (K1+K2), 12 lookups (K3), 12 lookup adjoints (K3b), pyramid adjoint (K1b+K2b); then the all-reduce
        lo, hi = shard_range(args.train_batch, world, rank)
        g = torch.Generator(device="cpu").manual_seed(99 + rank)
        # the rank's pairs go through in micro-batches (pyramid + fp32 gradient pyramid: 8.5 GB per pair in flight)
        TMB = max(1, min(args.train_micro, hi - lo))
        f1 = torch.randn(TMB, C, H, W, generator=g).to(dev).requires_grad_(True)
        f2 = torch.randn(TMB, C, H, W, generator=g).to(dev).requires_grad_(True)
        grid = E.coords_grid(TMB, H, W, device=dev)
        if smooth:
            # a RAFT-like trajectory: a smooth flow field of up to ~+-20 feature pixels (160 image pixels), approached
            # geometrically over the iterations, plus half a pixel of per-query noise
            flow = torch.nn.functional.interpolate(torch.randn(TMB, 2, 5, 8, generator=g) * 10, size=(H, W), mode="bicubic", align_corners=True).to(dev)
            coords = [(grid + flow * (1 - 0.5 ** (k + 1)) + torch.randn(TMB, 2, H, W, generator=g).to(dev) * 0.5).clamp_(-16, W + 16)
                      for k in range(ITERS)]
        else:
            coords = [(grid + torch.randn(TMB, 2, H, W, generator=g).to(dev) * (2 + k)).clamp_(-16, W + 16) for k in range(ITERS)]
        douts = [torch.randn(TMB, L * (2 * R + 1) ** 2, H, W, generator=g).to(dev) for _ in range(ITERS)]
        bucket = torch.randn(RAFT_PARAMS, device=dev)

        def train_step():
            for m in range(lo, hi, TMB):
                n = min(TMB, hi - m)
                f1.grad = f2.grad = None
                fn = E.MutliScalePairwise4DCorr(f1[:n], f2[:n], num_levels=L, corr_radius=R)
                loss = 0
                for k in range(ITERS):
                    loss = loss + (fn(coords[k][:n]) * douts[k][:n]).sum()
                loss.backward()
                del fn, loss
            if world > 1:
                import torch.distributed as dist

                dist.all_reduce(bucket)

        train_step()  # warm-up
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()
        nsteps = 2
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(nsteps):
            train_step()
        t1.record()
        torch.cuda.synchronize()
        (ms,) = max_over_ranks([t0.elapsed_time(t1) / nsteps], device=dev)
        per_pair = N * LEVEL_ELEMS * 2 + ITERS * 68.2e6 + ITERS * 145.7e6 + 2 * N * LEVEL_ELEMS * 4  # BASELINE.md §3 (fwd fp16 pyramid + bwd)
        add(c5_name, ms, int(per_pair * (hi - lo)),
            "RAFT 1080p correlation fwd+bwd training step, global batch sharded over the ranks, NCCL all-reduce(sum) of the "
            "RAFT-sized fp32 gradient bucket per step (BASELINE config 5); bytes = this rank's pairs; lookup coords = "
            + ("grid + a smooth flow field (+-20 px) approached over the 12 iterations + 0.5 px noise" if smooth else "grid + randn * (2 + k) for lookup k (windows scattered +-40 px)"),
            global_batch=args.train_batch, pairs_per_rank=hi - lo, micro_batch=TMB, pairs_per_s=args.train_batch / (ms / 1e3),
            ms_per_pair=ms / max(hi - lo, 1), allreduce_elems=RAFT_PARAMS if world > 1 else 0, scaling="strong", steps=nsteps)
        del f1, f2, coords, douts, bucket
        torch.cuda.empty_cache()

    if enabled("reference_ops_on_this_gpu"):
        # the honest GPU competitor: the unmodified reference class (cuBLAS bmm, avg_pool2d, grid_sample; TF32 off) on this GPU,
        # one 1080p pair at a time (its fp32 volume + pooled copies take ~7 GB per pair)
        cls = _import_reference_class()
        if cls is not None:
            old_tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
            torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False
            try:
                with torch.no_grad():
                    g = torch.Generator(device="cpu").manual_seed(5 + rank)
                    f1 = torch.randn(1, C, H, W, generator=g).to(dev)
                    f2 = torch.randn(1, C, H, W, generator=g).to(dev)
                    grid = E.coords_grid(1, H, W)
                    coords = [(grid + torch.randn(1, 2, H, W, generator=g) * (2 + k)).clamp_(-16, W + 16).to(dev) for k in range(ITERS)]

                    def ref_pair():
                        fn = cls(f1, f2, num_levels=L, corr_radius=R)
                        for k in range(ITERS):
                            fn(coords[k])

                    ms = timed(ref_pair, iters=5, warm=2)
                records.append({"name": "reference_ops_on_this_gpu", "ms": ms, "pairs_per_s": world * 1e3 / ms,
                                "note": "ezflow.similarity.MutliScalePairwise4DCorr (unmodified reference, PyTorch ops, TF32 off) on this GPU: "
                                        "build + 12 lookups of ONE 1080p pair per call; pairs/s = ranks / time"})
                del f1, f2, coords
            finally:
                torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old_tf32
            torch.cuda.empty_cache()

    if enabled("headline_sustained") and args.sustained_seconds > 0:
        # the headline step repeated for >= sustained_seconds of device time: the SM clock settles at the power cap
        with torch.no_grad():
            headline_step()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            headline_step()
            b.record()
            torch.cuda.synchronize()
            n = max(3, int(args.sustained_seconds * 1e3 / max(a.elapsed_time(b), 1e-3)) + 1)
            sampler = ClockSampler(dev.index or 0) if rank == 0 else None
            a.record()
            for _ in range(n):
                headline_step()
            b.record()
            torch.cuda.synchronize()
            clocks = sampler.stop() if sampler else None
            (ms,) = max_over_ranks([a.elapsed_time(b)], device=dev)
        pairs = world * args.pairs_per_gpu * n
        records.append({"name": "headline_sustained", "ms": ms, "steps": n, "pairs_per_s": pairs / (ms / 1e3), "clocks": clocks,
                        "note": "the headline step (build + 12 lookups, inputs resident) repeated back to back for >= %.1f s of device time" % args.sustained_seconds})
    return records


def run_b200(args):
    import torch

    import ezflow_b200 as E
    from ezflow_b200 import _cabi
    from ezflow_b200.similarity import pairwise as PW

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py --impl b200 needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _bind_to_gpu_numa_node(local_rank)  # pinned host buffers are then first-touched next to this GPU's PCIe root
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # keep stdout for the one JSON line
        dist.init_process_group("nccl", device_id=dev)
    E.set_pyramid_dtype(args.pyramid_dtype)
    lib = _cabi.lib()
    P, MB = args.pairs_per_gpu, args.micro
    assert P % MB == 0, "--pairs-per-gpu must be a multiple of --micro"

    g = torch.Generator(device="cpu").manual_seed(1234 + rank)
    # pinned host copies (e2e leg) and resident device copies (device-timed leg)
    h_f1 = torch.randn(P, C, H, W, generator=g).pin_memory()
    h_f2 = torch.randn(P, C, H, W, generator=g).pin_memory()
    grid = E.coords_grid(P, H, W)
    h_coords = [
        (grid + torch.randn(P, 2, H, W, generator=g) * (2 + k)).clamp_(-16, W + 16).pin_memory() for k in range(ITERS)
    ]
    d_f1, d_f2 = h_f1.to(dev), h_f2.to(dev)
    d_coords = [c.to(dev) for c in h_coords]

    # raw CUDA events recorded around every GEMM launch (roofline leg)
    ev_pairs = []

    def new_event():
        e = ctypes.c_void_p()
        _cabi.check(lib.ezb_event_create(ctypes.byref(e)))
        return e

    record_gemm = {"on": False}
    orig_build = PW._PyramidState.build

    def build_with_events(self, ev_gemm=None):
        if record_gemm["on"]:
            pair = (new_event(), new_event())
            ev_pairs.append(pair)
            return orig_build(self, ev_gemm=pair)
        return orig_build(self)

    PW._PyramidState.build = build_with_events

    def step(f1, f2, coords):
        last = None
        for m in range(0, P, MB):
            corr_fn = E.MutliScalePairwise4DCorr(f1[m : m + MB], f2[m : m + MB], num_levels=L, corr_radius=R)
            for k in range(ITERS):
                last = corr_fn(coords[k][m : m + MB])
            del corr_fn
        return last

    def barrier():
        if world > 1:
            import torch.distributed as dist

            dist.barrier()
        torch.cuda.synchronize()

    with torch.no_grad():
        for _ in range(args.warmup):
            step(d_f1, d_f2, d_coords)
        barrier()
        sampler = ClockSampler(local_rank) if rank == 0 else None
        record_gemm["on"] = True
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record()
        for _ in range(args.steps):
            step(d_f1, d_f2, d_coords)
        t1.record()
        barrier()
        record_gemm["on"] = False
        clocks = sampler.stop() if sampler else None
        ms = t0.elapsed_time(t1)

        gemm_ms = []
        for a, b in ev_pairs:
            v = ctypes.c_float()
            _cabi.check(lib.ezb_event_elapsed_ms(a, b, ctypes.byref(v)))
            gemm_ms.append(v.value)
            lib.ezb_event_destroy(a)
            lib.ezb_event_destroy(b)

        # ---- end-to-end leg: host buffers in, last lookup (the features the update block consumes) out
        e2e = None
        if not args.no_e2e:
            copy_s = torch.cuda.Stream(dev)
            out_host = torch.empty((P, L * (2 * R + 1) ** 2, H, W), dtype=torch.float32).pin_memory()
            bufs = [
                (torch.empty((MB, C, H, W), device=dev), torch.empty((MB, C, H, W), device=dev), [torch.empty((MB, 2, H, W), device=dev) for _ in range(ITERS)])
                for _ in range(2)
            ]

            def upload(m, slot):
                b1, b2, bc = bufs[slot]
                with torch.cuda.stream(copy_s):
                    b1.copy_(h_f1[m : m + MB], non_blocking=True)
                    b2.copy_(h_f2[m : m + MB], non_blocking=True)
                    for k in range(ITERS):
                        bc[k].copy_(h_coords[k][m : m + MB], non_blocking=True)
                    ev = torch.cuda.Event()
                    ev.record(copy_s)
                return ev

            d2h_s = torch.cuda.Stream(dev)

            def e2e_steps(nsteps):
                """`nsteps` steps as one stream of micro-batches: the H2D copy of micro-batch j+1, the kernels of
                micro-batch j and the D2H copy of micro-batch j-1 overlap (three streams, two buffer slots)."""
                cur = torch.cuda.current_stream(dev)
                nmb = P // MB
                total = nsteps * nmb
                ready = upload(0, 0)
                for j in range(total):
                    i, slot = j % nmb, j & 1
                    cur.wait_event(ready)
                    if j + 1 < total:
                        if j >= 1:
                            copy_s.wait_stream(cur)  # slot (j+1)&1 was last read by micro-batch j-1
                        ready = upload(((j + 1) % nmb) * MB, (j + 1) & 1)
                    b1, b2, bc = bufs[slot]
                    corr_fn = E.MutliScalePairwise4DCorr(b1, b2, num_levels=L, corr_radius=R)
                    last = None
                    for k in range(ITERS):
                        last = corr_fn(bc[k])
                    done = torch.cuda.Event()
                    done.record(cur)
                    d2h_s.wait_event(done)
                    with torch.cuda.stream(d2h_s):
                        out_host[i * MB : (i + 1) * MB].copy_(last, non_blocking=True)
                    last.record_stream(d2h_s)  # keep the allocator from reusing it before the copy has run
                    del corr_fn
                torch.cuda.synchronize()

            e2e_steps(max(1, min(args.warmup, 2)))
            barrier()
            w0 = time.perf_counter()
            e2e_steps(args.steps)
            barrier()
            e2e_s = time.perf_counter() - w0
            h2d = P * (2 * C * N * 4 + ITERS * 2 * N * 4)
            d2h = P * L * (2 * R + 1) ** 2 * N * 4
            e2e = (e2e_s, h2d, d2h)

    secondary = None
    if not args.no_secondary:
        PW._PyramidState.build = orig_build
        secondary = run_secondary(args, dev, world, rank, load_peaks(), lambda: step(d_f1, d_f2, d_coords))

    # max over ranks (device-timed step and end-to-end wall time)
    from ezflow_b200.sharding import max_over_ranks

    ms, e2e_s = max_over_ranks([ms, e2e[0] if e2e else 0.0], device=dev)
    if e2e:
        e2e = (e2e_s, e2e[1], e2e[2])

    if rank == 0:
        peaks = load_peaks()
        es = 2 if args.pyramid_dtype == "fp16" else 4
        total_pairs = world * P * args.steps
        value = total_pairs / (ms / 1e3)
        # dominant kernel: the tcgen05 GEMM with fused pyramid.  Algorithmic bytes per launch:
        # pyramid write N*42900*es + fp16 operand read 2*N*C*2, per pair, times MB pairs per launch.
        gemm_avg_ms = sum(gemm_ms) / max(len(gemm_ms), 1)
        bytes_per_launch = MB * (N * LEVEL_ELEMS * es + 2 * N * C * 2)
        flops_per_launch = MB * 2.0 * N * N * C
        achieved_gbs = bytes_per_launch / (gemm_avg_ms * 1e-3) / 1e9 if gemm_avg_ms > 0 else 0.0
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "gemm_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(f"{args.pyramid_dtype}_mb{MB}")
            except Exception:
                traffic = None
        line = {
            "metric": METRIC,
            "value": value,
            "unit": "pairs/s",
            "n_gpus": world,
            "steps": args.steps,
            "warmup": args.warmup,
            "ms_per_step": ms / args.steps,
            "higher_is_better": True,
            "scaling": "weak",
            "vs_baseline": None,
            "dtype": "f16",
            "data": "synthetic",
            "config": {
                "workload": "raft_1080p_corr_build_12_lookups",
                "feat": [C, H, W], "levels": L, "radius": R, "iters": ITERS,
                "pairs_per_gpu_per_step": P, "micro_batch": MB, "pyramid_storage": args.pyramid_dtype,
                "parallelism": f"batch-sharded x{world}, no collective",
                "l2": "inputs larger than L2 (531 MB of feature maps and a %.1f GB pyramid per step vs 126 MB L2)" % (P * N * LEVEL_ELEMS * es / 1e9),
            },
            "roofline": {
                "kernel": "corr_pyramid_gemm_kernel (K1+K2 fused)",
                "bound": "hbm",
                "achieved": achieved_gbs,
                "peak": peaks["hbm_gbs"],
                "unit": "GB/s",
                "frac": achieved_gbs / peaks["hbm_gbs"],
                "traffic": traffic,
                "peak_source": peaks["src"],
                "launch_ms": gemm_avg_ms,
                "launches_timed": len(gemm_ms),
                "tensor_tflops": flops_per_launch / (gemm_avg_ms * 1e-3) / 1e12 if gemm_avg_ms > 0 else 0.0,
                "tensor_frac_of_sustained_bf16": (flops_per_launch / (gemm_avg_ms * 1e-3) / 1e12) / peaks["tc_tflops"] if gemm_avg_ms > 0 else 0.0,
                "gemm_share_of_step": (sum(gemm_ms) / ms) if ms > 0 else None,
            },
            "clocks": clocks,
            "gpu_launches": args.steps * (P // MB) * (5 + ITERS),  # absmax, scales, convert, pool+pack, gemm + lookups (profiles/r02_launches.txt)
        }
        if e2e:
            line["e2e"] = {
                "value": world * P * args.steps / e2e[0],
                "unit": "pairs/s",
                "h2d_bytes_per_step": e2e[1],
                "d2h_bytes_per_step": e2e[2],
                "note": "pinned host fmaps+coords in, last lookup (B,324,H,W) out to pinned host; H2D, compute and D2H on three streams, double-buffered",
            }
        if world == 1 and not args.no_cpu_baseline:
            arm = ReferenceCpuArm()
            if arm.cls is not None:
                arm.step()  # warm-up (allocator, thread pool)
                dts = [arm.step() for _ in range(3)]
                line["cpu_baseline"] = {"value": 1.0 / min(dts), "unit": "pairs/s", "cores": arm.cores, "kind": "reference",
                                        "sample": arm.sample() + f"; 1 warm-up + best of 3 ({min(dts):.2f} s per pair)"}
            else:
                v, cores, sample = cpu_port_pairs_per_sec(args.cpu_sample_rows)
                line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample}
        if secondary is not None:
            line["secondary"] = secondary
        emit(line)
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


_JSON_FD = None


def _reserve_stdout():
    """stdout carries exactly one JSON line: file descriptor 1 is pointed at stderr for the rest of the process (NCCL and
    other native libraries print their banners with printf), and the JSON line is written to the saved descriptor."""
    global _JSON_FD
    sys.stdout.flush()
    _JSON_FD = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _JSON_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        sys.stdout.flush()
        os.write(_JSON_FD, data)


def main():
    args = parse_args()
    _reserve_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
