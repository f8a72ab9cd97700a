"""GPU parity at MODEL level: the reference's own RAFT / FlowNetC / PWCNet / DCVNet, built from the unmodified YAML configs,
run one forward on the GPU twice — once as shipped (the reference's PyTorch similarity ops) and once after
``ezflow_b200.install()`` (same weights, this repo's sm_100a kernels behind the same class names) — and the outputs are
compared directly.  north_star: RAFT final flow after 12 iterations within 0.01 px mean EPE; cost-volume models within 1e-3
relative.

Reference forwards exercised: ezflow/models/raft.py:96-129, models/flownet_c.py:86-90, decoder/pyramid.py:98-141,
models/dcvnet.py:62-84.  The reference is the pip-installed copy under baseline/_ref (git-ignored, made by
__graft_entry__.build(); it travels to the GPU box) — /root/reference does not exist there."""

import numpy as np
import pytest
import torch
import torch.nn.functional as F

import refshim
from gpu_util import dev

pytestmark = pytest.mark.gpu

EPE_TOL = 0.01  # px (north_star, RAFT)
REL_TOL = 1e-3  # relative Frobenius error of the final flow (cost-volume models)


def _smooth_pair(b, h, w, seed, shift=(3.5, -2.25)):
    """Low-pass-filtered noise and a copy shifted by a sub-pixel translation, so that the flow is non-trivial."""
    g = torch.Generator().manual_seed(seed)
    base = torch.randn(b, 3, h // 4 + 8, w // 4 + 8, generator=g)
    base = F.interpolate(base, size=(h + 32, w + 32), mode="bicubic", align_corners=False)
    img1 = base[:, :, 16 : 16 + h, 16 : 16 + w]
    ys, xs = torch.meshgrid(torch.arange(h, dtype=torch.float32), torch.arange(w, dtype=torch.float32), indexing="ij")
    gx = 2 * (xs + 16 + shift[0]) / (w + 31) - 1
    gy = 2 * (ys + 16 + shift[1]) / (h + 31) - 1
    img2 = F.grid_sample(base, torch.stack([gx, gy], -1)[None].expand(b, -1, -1, -1), align_corners=True)
    return img1.contiguous(), img2.contiguous()  # zero mean, unit-ish scale


@pytest.fixture()
def reference():
    ez = refshim.import_reference()
    if ez is None:
        pytest.skip("reference package not found (baseline/_ref is created by __graft_entry__.build())")
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = False  # both arms in true fp32 outside the similarity path
    torch.backends.cuda.matmul.allow_tf32 = False
    yield ez
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


def _build_both(name, yaml_name):
    import ezflow_b200 as E
    from ezflow.models import build_model

    cfg = refshim.reference_config_path(yaml_name)
    torch.manual_seed(0)
    ref_model = build_model(name, cfg_path=cfg, custom_cfg=True).to(dev()).eval()
    E.install()
    try:
        new_model = build_model(name, cfg_path=cfg, custom_cfg=True)
    finally:
        E.uninstall()
    missing = new_model.load_state_dict(ref_model.state_dict(), strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    return ref_model, new_model.to(dev()).eval()


def _forward_new(model, img1, img2):
    """The forward of the model built from B200 classes; install() stays active because RAFT / the PWC decoder read
    ``convex_upsample_flow`` / ``warp`` from their module globals at call time (ezflow/models/raft.py:127, decoder/pyramid.py:138)."""
    import ezflow_b200 as E

    E.install()
    try:
        with torch.no_grad():
            return model(img1, img2)
    finally:
        E.uninstall()


def _report(name, new, ref):
    epe = (new - ref).norm(dim=1).mean().item()
    rel = ((new - ref).norm() / ref.norm()).item()
    mag = ref.norm(dim=1).mean().item()
    print(f"{name}: mean |flow| {mag:.3f} px, EPE(b200 classes, reference) = {epe:.5f} px, rel_fro = {rel:.2e}")
    return epe, rel


def test_raft_yaml_forward_epe(reference, golden):
    import ezflow_b200 as E

    ref_model, new_model = _build_both("RAFT", "raft.yaml")
    assert new_model.similarity_fn is E.MutliScalePairwise4DCorr and ref_model.similarity_fn is not E.MutliScalePairwise4DCorr
    g = golden("raft_e2e")  # BASELINE config 1: the 368x496 8-bit image pair of tools/make_raft_e2e.py
    img1, img2 = (torch.from_numpy(g[k]).float().to(dev()) for k in ("img1", "img2"))
    with torch.no_grad():
        ref = ref_model(img1, img2)["flow_upsampled"]
    new = _forward_new(new_model, img1, img2)["flow_upsampled"]
    gold = torch.from_numpy(g["flow"].astype(np.float32)).to(dev())  # the reference's own CPU run of the same model
    epe_cpu_ref = (ref - gold).norm(dim=1).mean().item()
    epe_cpu_new = (new - gold).norm(dim=1).mean().item()
    print(f"RAFT: EPE(reference on this GPU, reference CPU run) = {epe_cpu_ref:.5f} px, "
          f"EPE(b200 classes, reference CPU run) = {epe_cpu_new:.5f} px")
    # direct: this repo's classes on the GPU against the reference's own CPU output (includes CPU-vs-GPU convolution rounding)
    assert epe_cpu_new <= EPE_TOL, f"mean EPE vs the reference's CPU run {epe_cpu_new:.4f} px > {EPE_TOL}"
    assert new.shape == ref.shape and torch.isfinite(new).all()
    epe, _ = _report("RAFT raft.yaml 368x496, 12 iterations", new, ref)
    assert epe <= EPE_TOL, f"mean EPE {epe:.4f} px > {EPE_TOL}"


@pytest.mark.parametrize("mode", ["tc", "exact"])
def test_flownetc_yaml_forward(reference, mode):
    import ezflow_b200 as E

    ref_model, new_model = _build_both("FlowNetC", "flownet_c.yaml")
    assert isinstance(new_model.correlation_layer, E.IterSpatialCorrelationSampler)
    img1, img2 = (t.to(dev()) for t in _smooth_pair(2, 384, 512, seed=7))
    old = E.get_local_corr_mode()
    E.set_local_corr_mode(mode)
    try:
        with torch.no_grad():
            ref = ref_model(img1, img2)["flow_upsampled"]
        new = _forward_new(new_model, img1, img2)["flow_upsampled"]
        _, rel = _report(f"FlowNetC flownet_c.yaml 384x512 ({mode})", new, ref)
        assert rel <= REL_TOL
        if mode == "tc":  # + LeakyReLU folded into the sampler's epilogue (fuse_model)
            assert E.fuse_model(new_model) == 1
            fused = _forward_new(new_model, img1, img2)["flow_upsampled"]
            _, rel = _report("FlowNetC, fused activation", fused, ref)
            assert rel <= REL_TOL
    finally:
        E.set_local_corr_mode(old)


@pytest.mark.parametrize("mode", ["tc", "exact"])
def test_pwcnet_yaml_forward(reference, mode):
    import ezflow_b200 as E

    ref_model, new_model = _build_both("PWCNet", "pwcnet.yaml")
    assert isinstance(new_model.decoder.correlation_layer, E.IterSpatialCorrelationSampler)
    img1, img2 = (t.to(dev()) for t in _smooth_pair(2, 448, 1024, seed=11))
    old = E.get_local_corr_mode()
    E.set_local_corr_mode(mode)
    try:
        with torch.no_grad():
            ref = ref_model(img1, img2)["flow_upsampled"]
        new = _forward_new(new_model, img1, img2)["flow_upsampled"]  # lazy warp -> sampler fusion active
        _, rel = _report(f"PWCNet pwcnet.yaml 448x1024 ({mode})", new, ref)
        assert rel <= REL_TOL
        if mode == "tc":
            assert E.fuse_model(new_model) == 1
            fused = _forward_new(new_model, img1, img2)["flow_upsampled"]
            _, rel = _report("PWCNet, fused warp + activation", fused, ref)
            assert rel <= REL_TOL
    finally:
        E.set_local_corr_mode(old)


@pytest.mark.parametrize("mode", ["tc", "exact"])
def test_dcvnet_yaml_forward(reference, mode):
    import ezflow_b200 as E

    ref_model, new_model = _build_both("DCVNet", "dcvnet.yaml")
    assert isinstance(new_model.cost_volume_list, E.MatryoshkaDilatedCostVolumeList)
    img1, img2 = (t.to(dev()) for t in _smooth_pair(1, 384, 768, seed=5))
    old = E.get_local_corr_mode()
    E.set_local_corr_mode(mode)
    try:
        with torch.no_grad():
            ref = ref_model(img1, img2)["flow_upsampled"]
        new = _forward_new(new_model, img1, img2)["flow_upsampled"]
        _, rel = _report(f"DCVNet dcvnet.yaml 384x768 ({mode})", new, ref)
        assert rel <= REL_TOL
    finally:
        E.set_local_corr_mode(old)


def test_raft_yaml_training_step_parameter_gradients(reference):
    """BASELINE config 5 through the model: one training-mode forward + backward of the reference's RAFT (raft.yaml, 12
    iterations, sequence-loss-like weights) with the reference's PyTorch similarity ops and with this repo's classes (same
    weights): the gradients that reach the feature-encoder parameters pass through the whole adjoint path — deferred lookup
    adjoints in one launch, support ranges, tf32 pyramid-adjoint GEMMs — and must agree.  Reference: engine/trainer.py:269-299
    (loss.backward()), models/raft.py:96-129."""
    import ezflow_b200 as E

    ref_model, new_model = _build_both("RAFT", "raft.yaml")
    ref_model.train()
    new_model.train()
    img1, img2 = (t.to(dev()) for t in _smooth_pair(2, 128, 160, seed=3))
    target = torch.zeros(2, 2, 128, 160, device=dev())
    target[:, 0], target[:, 1] = 3.5, -2.25

    def loss_of(preds):
        n = len(preds)
        return sum(0.8 ** (n - 1 - i) * (p - target).abs().mean() for i, p in enumerate(preds))

    loss_ref = loss_of(ref_model(img1, img2)["flow_preds"])
    loss_ref.backward()
    E.install()
    try:
        loss_new = loss_of(new_model(img1, img2)["flow_preds"])
        loss_new.backward()
    finally:
        E.uninstall()
    assert abs(loss_new.item() - loss_ref.item()) <= 1e-3 * abs(loss_ref.item()), (loss_new.item(), loss_ref.item())
    worst, checked, at_conv_out = 0.0, 0, None
    ref_params = dict(ref_model.named_parameters())
    gmax = max(p.grad.norm().item() for n, p in ref_params.items() if n.startswith("fnet.") and p.grad is not None)
    for name, p in new_model.named_parameters():
        if not name.startswith("fnet.") or p.grad is None:
            continue  # the feature encoder: every gradient it gets comes through the correlation pyramid's adjoint
        gr = ref_params[name].grad
        assert gr is not None and torch.isfinite(p.grad).all()
        if gr.norm().item() <= 1e-6 * gmax:
            # a convolution bias in front of an instance norm: its gradient is zero analytically (1e-9 of rounding noise here)
            assert p.grad.norm().item() <= 1e-5 * gmax
            continue
        rel = ((p.grad - gr).norm() / gr.norm()).item()
        worst = max(worst, rel)
        checked += 1
        if name == "fnet.conv_out.weight":
            at_conv_out = rel
    print(f"RAFT training step 128x160: loss {loss_new.item():.6f} vs {loss_ref.item():.6f}; {checked} feature-encoder parameter "
          f"tensors, worst relative gradient error {worst:.2e} (at the encoder's output convolution {at_conv_out:.2e})")
    # measured: 1.3e-3 at the output convolution (fp16 pyramid forward, tf32 adjoint GEMMs), 5.6e-3 further up (the instance
    # norms' backward subtracts two nearly equal terms and amplifies it)
    assert checked >= 10 and at_conv_out is not None
    assert at_conv_out <= 5e-3, f"gradient at the encoder's output convolution differs by {at_conv_out:.2e} relative"
    assert worst <= 2e-2, f"feature-encoder gradients differ by {worst:.2e} relative"
    E.release_gradient_pool()


def _capture_similarity_inputs(module, store, first_only=False):
    """Forward pre-hook: keeps every floating-point tensor entering the similarity module (lists included) and asks autograd to
    retain its gradient — the quantity this repo's adjoint kernels produce, before any parameter reduction."""

    def hook(_mod, args):
        for a in (args[:1] if first_only else args):
            for t in (a if isinstance(a, (list, tuple)) else [a]):
                if torch.is_tensor(t) and t.requires_grad:
                    t.retain_grad()
                    store.append(t)

    return module.register_forward_pre_hook(hook)


@pytest.mark.parametrize("which", ["FlowNetC", "PWCNet", "DCVNet"])
@pytest.mark.parametrize("mode", ["tc", "exact"])
def test_cost_volume_models_training_step_gradients(reference, which, mode):
    """Training-mode forward + backward of the reference's FlowNetC / PWCNet / DCVNet with the reference's sampler ops and with
    this repo's classes (same weights): the loss, and the gradients arriving at every input of the similarity module (feature
    maps; for PWCNet the warped features of all five levels) — what K4b (tensor-core band GEMMs in 'tc' mode, register-tiled /
    tiled fp32 in 'exact'), the Matryoshka adjoint incl. the stride-4 scale (DCVNet) and the chain behind them produce.
    The tolerance is measured, not guessed: the same quantities of the reference itself when its own similarity output is
    perturbed by what this repo's forward differs from it (see `eps` below); parameter gradients likewise.
    Reference: engine/trainer.py:269-299."""
    import ezflow_b200 as E

    yaml_name, size, seed = {"FlowNetC": ("flownet_c.yaml", (256, 320), 7), "PWCNet": ("pwcnet.yaml", (256, 448), 11),
                             "DCVNet": ("dcvnet.yaml", (256, 384), 5)}[which]
    ref_model, new_model = _build_both(which, yaml_name)
    ref_model.train()
    new_model.train()
    sim = {"FlowNetC": lambda m: m.correlation_layer, "PWCNet": lambda m: m.decoder.correlation_layer, "DCVNet": lambda m: m.cost_volume_list}[which]
    img1, img2 = (t.to(dev()) for t in _smooth_pair(2, size[0], size[1], seed=seed))

    def loss_of(out):
        return sum(p.abs().mean() + 0.1 * (p * p).mean() for p in out["flow_preds"])

    old = E.get_local_corr_mode()
    E.set_local_corr_mode(mode)
    try:
        # The yardstick: the reference against itself when its own similarity output is perturbed at the level this repo's
        # forward differs from it (fp32 rounding in 'exact' mode, 3e-4 of fp16 operands in 'tc' mode).  A deep network in
        # training mode (batch statistics, LeakyReLU kinks, cancelling sums in bias gradients) amplifies such a perturbation,
        # and cuDNN's backward kernels add run-to-run differences of their own; an adjoint that is right stays within a small
        # multiple of that, a wrong one does not.
        # 'exact': 2e-6 — the fp32 kernels differ from the reference's op sequence by summation order (K4 'exact' is tested
        # to 2e-5 against the oracle), and the yardstick has to cover cuDNN's own run-to-run differences, which at 3e-7 it did
        # not (PWCNet's bias gradients moved by 0.9-2.3e-2 beyond it from one run to the next on the same box).
        eps = 3e-4 if mode == "tc" else 2e-6
        gen = torch.Generator(device=dev()).manual_seed(123)

        def perturb(_mod, _args, out):
            return out + eps * out.abs().mean() * torch.randn(out.shape, device=out.device, generator=gen)

        pert_in = []
        h = _capture_similarity_inputs(sim(ref_model), pert_in, first_only=which == "PWCNet")
        h2 = sim(ref_model).register_forward_hook(perturb)
        torch.manual_seed(1)
        loss_of(ref_model(img1, img2)).backward()
        h.remove()
        h2.remove()
        first = {n: p.grad.clone() for n, p in ref_model.named_parameters() if p.grad is not None}
        first_in = [t.grad.clone() for t in pert_in]
        ref_model.zero_grad(set_to_none=True)
        ref_in = []
        # PWCNet: this repo's sampler receives a lazy-warp handle as its second input (the warped tensor is never materialised)
        h = _capture_similarity_inputs(sim(ref_model), ref_in, first_only=which == "PWCNet")
        torch.manual_seed(1)
        loss_ref = loss_of(ref_model(img1, img2))
        loss_ref.backward()
        h.remove()
        noise = {n: ((p.grad - first[n]).norm() / first[n].norm().clamp_min(1e-30)).item() for n, p in ref_model.named_parameters() if p.grad is not None}
        # the same yardstick over all parameters at once (dominated by the large gradients, whose relative noise is stable)
        noise_all = (sum(((p.grad - first[n]) ** 2).sum().item() for n, p in ref_model.named_parameters() if p.grad is not None)
                     / sum((first[n] ** 2).sum().item() for n in first)) ** 0.5
        noise_in = max(((a - t.grad).norm() / t.grad.norm().clamp_min(1e-30)).item() for a, t in zip(first_in, ref_in))
        new_in = []
        h = _capture_similarity_inputs(sim(new_model), new_in, first_only=which == "PWCNet")
        E.install()
        try:
            torch.manual_seed(1)
            loss_new = loss_of(new_model(img1, img2))
            loss_new.backward()
        finally:
            E.uninstall()
            h.remove()
    finally:
        E.set_local_corr_mode(old)
    tol_loss = 1e-3 if mode == "tc" else 1e-5
    assert abs(loss_new.item() - loss_ref.item()) <= tol_loss * abs(loss_ref.item()), (loss_new.item(), loss_ref.item())
    assert len(new_in) == len(ref_in) and len(ref_in) >= 2
    worst_in = 0.0
    for i, (a, b) in enumerate(zip(new_in, ref_in)):
        assert a.grad is not None and b.grad is not None and a.grad.shape == b.grad.shape
        worst_in = max(worst_in, ((a.grad - b.grad).norm() / b.grad.norm().clamp_min(1e-30)).item())
    ref_params = dict(ref_model.named_parameters())
    gmax = max(p.grad.norm().item() for p in ref_params.values() if p.grad is not None)
    worst, worst_name, checked = 0.0, "", 0
    err2, ref2 = 0.0, 0.0
    for name, p in new_model.named_parameters():
        gr = ref_params[name].grad
        if gr is None or p.grad is None:
            assert gr is None and p.grad is None, name
            continue
        assert torch.isfinite(p.grad).all(), name
        err2 += ((p.grad - gr) ** 2).sum().item()
        ref2 += (gr ** 2).sum().item()
        if gr.norm().item() <= 1e-2 * gmax:
            continue  # analytically zero (a bias in front of a normalisation) or small: judged with all tensors together below
        rel = ((p.grad - gr).norm() / gr.norm()).item() - 4.0 * noise.get(name, 0.0)
        if rel > worst:
            worst, worst_name = rel, name
        checked += 1
    print(f"\n{which} training step {size} ({mode}): loss {loss_new.item():.6f} vs {loss_ref.item():.6f}; gradients at the {len(ref_in)} "
          f"similarity inputs within {worst_in:.2e} (the reference under a {eps:.0e} perturbation of its own similarity output: {noise_in:.2e}); "
          f"{checked} parameter tensors, worst relative difference beyond 4x that yardstick {worst:.2e} ({worst_name})")
    assert worst_in <= 4.0 * noise_in + 1e-5, f"gradients at the similarity module's inputs differ by {worst_in:.2e} relative (yardstick {noise_in:.2e})"
    # Parameter gradients.  Tight bound on all of them together: 4x what the reference itself moves by under the perturbation.
    # Per tensor only a gross-error guard: a single small bias gradient (a cancelling sum over a whole feature map) moves by
    # several percent (up to 13 % seen) from one run to the next on the same box — one perturbed run is one sample of that
    # noise, and a per-tensor bound of 2e-2 beyond it failed 2 runs in 12 (PWCNet's deconvolution biases) with correct kernels.
    # So the per-tensor guard covers the tensors with a significant gradient (>= 1 % of the largest); a wrong adjoint shows up
    # as an O(1) error in every tensor upstream of the similarity module.
    err_all = (err2 / max(ref2, 1e-60)) ** 0.5
    print(f"{which} ({mode}): all parameter gradients together differ by {err_all:.2e} relative (reference under the perturbation: {noise_all:.2e})")
    assert err_all <= 4.0 * noise_all + 1e-5, f"parameter gradients differ by {err_all:.2e} relative over all tensors (yardstick {noise_all:.2e})"
    assert checked >= 5 and worst <= 0.25, f"{worst_name}: parameter gradients differ by {worst:.2e} relative beyond the yardstick"
