"""GPU: re-entrancy.  The reference wraps models in nn.DataParallel for evaluation (ezflow/engine/eval.py:306,317): one Python
thread per device calls the similarity module concurrently.  The library keeps no global mutable state; this test runs the
RAFT path and the local correlation from several threads at once (on every visible device) and checks the results against a
single-threaded run."""

import threading

import pytest
import torch

pytestmark = pytest.mark.gpu


def _work(device, seed):
    import ezflow_b200 as E

    g = torch.Generator(device="cpu").manual_seed(seed)
    dev = torch.device("cuda", device)
    f1 = torch.randn(2, 64, 24, 40, generator=g).to(dev)
    f2 = torch.randn(2, 64, 24, 40, generator=g).to(dev)
    c = E.coords_grid(2, 24, 40, device=dev) + torch.randn(2, 2, 24, 40, generator=g).to(dev) * 3
    with torch.cuda.device(dev):
        out = E.MutliScalePairwise4DCorr(f1, f2, num_levels=3, corr_radius=3)(c)
        loc = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=9)(f1, f2)
    return out.cpu(), loc.cpu()


def test_concurrent_calls_from_threads():
    n_dev = torch.cuda.device_count()
    jobs = [(d, 100 + i) for i, d in enumerate(list(range(n_dev)) * (4 if n_dev == 1 else 2))]
    expect = [_work(d, s) for d, s in jobs]
    got = [None] * len(jobs)
    errs = []

    def run(i, d, s):
        try:
            for _ in range(5):
                got[i] = _work(d, s)
        except Exception as e:  # pragma: no cover
            errs.append(e)

    ts = [threading.Thread(target=run, args=(i, d, s)) for i, (d, s) in enumerate(jobs)]
    for t in ts:
        t.start()
    for t in ts:
        t.join()
    assert not errs, errs
    for (a, b), (x, y) in zip(expect, got):
        assert torch.equal(a, x) and torch.equal(b, y)
