#!/usr/bin/env python
"""BASELINE config 5: RAFT 1080x1920 all-pairs correlation fwd+bwd "training step", batch 64 sharded over the GPUs of one
node, followed by an NCCL all-reduce (sum) of a RAFT-sized gradient bucket (5,257,536 fp32 = the parameter count of
configs/models/raft.yaml), standing in for DDP (reference ezflow/engine/trainer.py:283,785).

    python tools/bench_train.py                                   # 1 GPU
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/bench_train.py --steps 2

Per pair: pyramid build (K1+K2), 12 lookups (K3), 12 lookup adjoints (K3b), pyramid adjoint (K1b+K2b).  Prints one JSON line
(pairs/s over all ranks, max-over-ranks device time).  The similarity path has no parameters: the all-reduce moves the
model's gradient bucket only; df1/df2 are per-sample and stay local."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ezflow_b200 as E
from ezflow_b200.sharding import max_over_ranks, shard_range

H, W, C, L, R, ITERS = 135, 240, 256, 4, 4, 12
RAFT_PARAMS = 5_257_536

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=64)
ap.add_argument("--steps", type=int, default=2)
ap.add_argument("--warmup", type=int, default=1)
args = ap.parse_args()
world, rank, local = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")), int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    import torch.distributed as dist

    dist.init_process_group("nccl", device_id=dev)
lo, hi = shard_range(args.batch, world, rank)
g = torch.Generator(device="cpu").manual_seed(99 + rank)
# one pair's worth of distinct inputs is re-used for the rank's pairs (values do not affect the timing)
f1 = torch.randn(1, C, H, W, generator=g).to(dev).requires_grad_(True)
f2 = torch.randn(1, C, H, W, generator=g).to(dev).requires_grad_(True)
grid = E.coords_grid(1, H, W, device=dev)
coords = [(grid + torch.randn(1, 2, H, W, generator=g).to(dev) * (2 + k)).clamp_(-16, W + 16) for k in range(ITERS)]
douts = [torch.randn(1, L * (2 * R + 1) ** 2, H, W, generator=g).to(dev) for _ in range(ITERS)]
bucket = torch.randn(RAFT_PARAMS, device=dev)


def step():
    for _ in range(lo, hi):
        f1.grad = f2.grad = None
        fn = E.MutliScalePairwise4DCorr(f1, f2, num_levels=L, corr_radius=R)
        loss = 0
        for k in range(ITERS):
            loss = loss + (fn(coords[k]) * douts[k]).sum()
        loss.backward()
        del fn, loss
    if world > 1:
        dist.all_reduce(bucket)


for _ in range(args.warmup):
    step()
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(args.steps):
    step()
t1.record()
torch.cuda.synchronize()
(ms,) = max_over_ranks([t0.elapsed_time(t1)], device=dev)
if rank == 0:
    print(json.dumps({"workload": "raft_1080p_corr_fwd_bwd_train_step", "n_gpus": world, "global_batch": args.batch,
                      "steps": args.steps, "ms_per_step": ms / args.steps, "pairs_per_s": args.batch * args.steps / (ms / 1e3),
                      "ms_per_pair_per_gpu": ms / args.steps / (hi - lo), "allreduce_elems": RAFT_PARAMS if world > 1 else 0,
                      "scaling": "strong"}))
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
