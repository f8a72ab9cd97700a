#!/usr/bin/env python
"""Headline benchmark: RAFT 1080p correlation-pyramid build + 12 lookups, image pairs / second.

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a path
    python bench.py --impl reference --gpus N ...            # CPU port of the reference path (host cores)

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
    ap.add_argument("--cpu-sample-rows", type=int, default=4, help="query rows (of 135) per worker in the CPU sample")
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


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times = []
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
        "config": {"workload": "raft_1080p_corr_build_12_lookups", "feat": [C, H, W], "levels": L, "radius": R, "iters": ITERS},
        "cpu_baseline": {"value": value, "unit": "pairs/s", "cores": workers, "kind": "port", "sample": sample},
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
            "gpu_launches": args.steps * (P // MB) * (4 + (L - 1) + 1 + ITERS),  # absmax, scales, convert, pools, pack, gemm + lookups
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
            v, cores, sample = cpu_port_pairs_per_sec(args.cpu_sample_rows)
            line["cpu_baseline"] = {"value": v, "unit": "pairs/s", "cores": cores, "kind": "port", "sample": sample}
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
