"""N>1 host logic on CPU: world_size-2 gloo process group (no GPU).  Pairs are sharded along the batch dimension with no
data-path collective; the only collectives are the barrier and the max/sum of per-rank scalars bench.py reports."""

import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from ezflow_b200.sharding import max_over_ranks, shard_range, sum_over_ranks

        lo, hi = shard_range(13, world, rank)
        dist.barrier()
        ms = 10.0 + 5.0 * rank  # rank 1 is slower
        (mx,) = max_over_ranks([ms])
        (total,) = sum_over_ranks([hi - lo])
        # config 5 (bench.py): every rank takes its block of the global batch, then ONE all-reduce(sum) of the gradient bucket
        blo, bhi = shard_range(64, world, rank)
        bucket = torch.full((1024,), float(bhi - blo))
        dist.all_reduce(bucket)
        assert bucket.eq(64.0).all()
        q.put((rank, lo, hi, mx, total))
    finally:
        dist.destroy_process_group()


def test_world_size_2_sharding_and_reductions():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 7), (7, 13)]      # contiguous, balanced, covering
    assert all(r[3] == 15.0 for r in res)                         # max over ranks of the step time
    assert all(r[4] == 13.0 for r in res)                         # every pair is owned exactly once


def test_shard_range_properties():
    from ezflow_b200.sharding import shard_range

    for n in (0, 1, 7, 64, 65):
        for world in (1, 2, 4, 8):
            blocks = [shard_range(n, world, r) for r in range(world)]
            assert blocks[0][0] == 0 and blocks[-1][1] == n
            assert all(blocks[i][1] == blocks[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in blocks]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_single_process_reductions_are_identity():
    from ezflow_b200.sharding import max_over_ranks, sum_over_ranks

    assert max_over_ranks([1.5, 2.0]) == [1.5, 2.0] and sum_over_ranks([3]) == [3.0]
