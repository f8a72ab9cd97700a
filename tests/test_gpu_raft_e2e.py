"""GPU parity, end to end: RAFT (configs/models/raft.yaml, seeded random init, 368x496, 12 iterations) with this repo's
similarity path vs the reference — north_star: final flow within 0.01 px mean EPE.

The reference cannot travel to the GPU box, so tools/make_raft_e2e.py (run where /root/reference exists) leaves
  * tests/golden/raft_e2e.npz       the image pair and the reference's own final flow (CPU, fp32), and
  * oracle/_ref/raft_cfg1_parts.pt  TorchScript traces of the reference's non-similarity parts (encoders, update block +
                                    convex upsampling) holding the same weights.
Here the 12-iteration loop (reference ezflow/models/raft.py:96-129) is run on the GPU around
  (a) a plain PyTorch fp32 restatement of the similarity path (the floating-point reference for this kernel), and
  (b) the CUDA path through the drop-in class,
so that conv-backend differences between the CPU run and the GPU cancel in (a) vs (b)."""

import os

import numpy as np
import pytest
import torch

from gpu_util import TorchPairwiseRef, dev

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PARTS = os.path.join(ROOT, "oracle", "_ref", "raft_cfg1_parts.pt")
EPE_TOL = 0.01  # px, mean end-point error of the final upsampled flow (north_star)


def _run(parts, corr_cls, img1, img2, levels, radius, iters):
    f1, f2, net, inp = parts.front(img1, img2)
    fn = corr_cls(f1.contiguous(), f2.contiguous(), num_levels=levels, corr_radius=radius)
    b, _, h, w = f1.shape
    ys, xs = torch.meshgrid(torch.arange(h, device=f1.device), torch.arange(w, device=f1.device), indexing="ij")
    c0 = torch.stack([xs, ys], dim=0).float()[None].repeat(b, 1, 1, 1)
    c1 = c0.clone()
    up = None
    for _ in range(iters):
        net, c1, up = parts.step(net, inp, fn(c1), c0, c1)
    return up


@pytest.mark.skipif(not os.path.exists(PARTS), reason="oracle/_ref/raft_cfg1_parts.pt missing (python tools/make_raft_e2e.py)")
def test_raft_cfg1_final_flow_epe(golden):
    import ezflow_b200 as E

    g = golden("raft_e2e")
    L, r, iters = int(g["levels"]), int(g["radius"]), int(g["iters"])
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False  # the reference ran in true fp32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        parts = torch.jit.load(PARTS, map_location=dev()).eval()
        img1 = torch.from_numpy(g["img1"]).to(dev()).float()
        img2 = torch.from_numpy(g["img2"]).to(dev()).float()
        with torch.no_grad():
            flow_ref = _run(parts, TorchPairwiseRef, img1, img2, L, r, iters)
            flow_new = _run(parts, E.MutliScalePairwise4DCorr, img1, img2, L, r, iters)
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
    assert tuple(flow_new.shape) == g["flow"].shape and torch.isfinite(flow_new).all()
    epe = (flow_new - flow_ref).norm(dim=1).mean().item()
    gold = torch.from_numpy(g["flow"].astype(np.float32)).to(dev())
    epe_gold_ref = (flow_ref - gold).norm(dim=1).mean().item()
    epe_gold_new = (flow_new - gold).norm(dim=1).mean().item()
    mag = gold.norm(dim=1).mean().item()
    print(f"RAFT cfg1: mean |flow| {mag:.2f} px; EPE(cuda path, torch fp32 similarity) = {epe:.4f} px; "
          f"vs reference CPU run: torch-ref {epe_gold_ref:.4f}, cuda path {epe_gold_new:.4f}")
    assert epe <= EPE_TOL, f"mean EPE {epe:.4f} px > {EPE_TOL}"
    # the GPU re-run of the reference pipeline itself differs from its CPU run only by conv/GEMM backend rounding;
    # the CUDA path must not be further from the reference's own output than that plus the tolerance
    assert epe_gold_new <= epe_gold_ref + EPE_TOL
    # (The direct comparison with the reference's own forward — same GPU, and against its CPU run — is asserted at 0.01 px in
    # tests/test_gpu_models_dropin.py::test_raft_yaml_forward_epe, where the unmodified reference model itself runs.  Here the
    # TorchScript-traced parts differ from the CPU run by ~0.05 px on their own, with either similarity path.)
