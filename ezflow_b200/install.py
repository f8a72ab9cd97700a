"""Installs the B200 similarity classes into a live ``ezflow`` so its unmodified YAML configs
(configs/models/raft.yaml, flownet_c.yaml, pwcnet.yaml, dcvnet.yaml) pick them up.

The reference's ``Registry`` asserts on duplicate names (ezflow/utils/registry.py:23-25), so the
entries are overwritten in ``_obj_map``.  FlowNetC and PWC-Net bypass the registry and import the
sampler by name (ezflow/models/flownet_c.py:9, ezflow/decoder/pyramid.py:6), and
``learnable_cost`` binds a module-global alias (ezflow/similarity/learnable_cost.py:6-9); those
globals are rebound too, as are ``warp`` in the PWC decoder (ezflow/decoder/pyramid.py:138) and
``convex_upsample_flow`` in RAFT and DCVNet's decoder (ezflow/models/raft.py:127, decoder/dilated_flow_stack_filter.py:219).
All of them are read at model-construction / call time, so ``install()`` must run before ``build_model``.
"""

import importlib

from . import similarity as S
from . import utils as U

_REGISTRY_NAMES = (
    "MutliScalePairwise4DCorr",
    "IterSpatialCorrelationSampler",
    "CorrelationLayer",
    "MatryoshkaDilatedCostVolume",
    "MatryoshkaDilatedCostVolumeList",
)

_saved = {}


def _rebind(modname, attr, value):
    mod = importlib.import_module(modname)
    key = (modname, attr)
    if key not in _saved:
        _saved[key] = getattr(mod, attr, None)
    setattr(mod, attr, value)


def install(patch_warp=True, patch_upsample=True, fuse_warp=True):
    """Returns the list of (module, attribute) pairs that were rebound.

    fuse_warp: the PWC decoder's ``warp`` (whose only consumer is the correlation sampler, ezflow/decoder/pyramid.py:138-141)
    returns a ``WarpedFeatures`` handle, so the sampler evaluates the warp inside its own kernel."""
    ez_sim = importlib.import_module("ezflow.similarity")
    reg = ez_sim.SIMILARITY_REGISTRY
    for name in _REGISTRY_NAMES:
        key = ("<registry>", name)
        if key not in _saved:
            _saved[key] = reg._obj_map.get(name)
        reg._obj_map[name] = getattr(S, name)
        _rebind("ezflow.similarity", name, getattr(S, name))
    for name in ("MutliScalePairwise4DCorr", "IterSpatialCorrelationSampler", "CorrelationLayer"):
        _rebind("ezflow.similarity.correlation", name, getattr(S, name))
    _rebind("ezflow.similarity.correlation.sampler", "iter_spatial_correlation_sample", S.iter_spatial_correlation_sample)
    for name in ("MatryoshkaDilatedCostVolume", "MatryoshkaDilatedCostVolumeList"):
        _rebind("ezflow.similarity.learnable_cost", name, getattr(S, name))
    # by-name imports that bypass the registry
    _rebind("ezflow.models.flownet_c", "SpatialCorrelationSampler", S.IterSpatialCorrelationSampler)
    _rebind("ezflow.decoder.pyramid", "SpatialCorrelationSampler", S.IterSpatialCorrelationSampler)
    _rebind("ezflow.similarity.learnable_cost", "SpatialCorrelationSampler", S.IterSpatialCorrelationSampler)
    if patch_warp:
        _rebind("ezflow.decoder.pyramid", "warp", S.lazy_warp if fuse_warp else U.warp)
        _rebind("ezflow.utils.warp", "warp", U.warp)
        _rebind("ezflow.utils", "warp", U.warp)
    # VCN's inline correlation (a method, ezflow/models/vcn.py:162-206) and DICL's volume assembly
    # (LearnableMatchingCost.forward, ezflow/similarity/learnable_cost.py:157-231; the class keeps its matching network)
    from .similarity import volume as VOL

    try:
        vcn_mod = importlib.import_module("ezflow.models.vcn")
        key = ("ezflow.models.vcn", "VCN._corr_fn")
        if key not in _saved:
            _saved[key] = vcn_mod.VCN._corr_fn
        vcn_mod.VCN._corr_fn = lambda self, features1, features2, max_disp, factorization=1: VOL.vcn_corr(features1, features2, max_disp, factorization)
        lc_mod = importlib.import_module("ezflow.similarity.learnable_cost")
        key = ("ezflow.similarity.learnable_cost", "LearnableMatchingCost.forward")
        if key not in _saved:
            _saved[key] = lc_mod.LearnableMatchingCost.forward
        lc_mod.LearnableMatchingCost.forward = VOL.learnable_matching_cost_forward
    except (ImportError, AttributeError):
        pass
    if patch_upsample:
        # convex upsampling of every iteration's flow (ezflow/models/raft.py:9,127; decoder/dilated_flow_stack_filter.py:8,219)
        for modname in ("ezflow.utils.resampling", "ezflow.utils", "ezflow.models.raft", "ezflow.decoder.dilated_flow_stack_filter"):
            _rebind(modname, "convex_upsample_flow", U.convex_upsample_flow)
    return [k for k in _saved if k[0] != "<registry>"]


def fuse_model(model):
    """Post-build surgery: wherever a B200 correlation sampler is directly followed by a LeakyReLU — FlowNetC
    (``corr_activation``, ezflow/models/flownet_c.py:42,86-90) and the PWC decoder (``leaky_relu``,
    ezflow/decoder/pyramid.py:49,98-102), neither of which uses that activation anywhere else — the activation is folded into
    the sampler's epilogue and replaced by ``nn.Identity``; in RAFT the lookup is fused with the update block's first 1x1
    convolution.  Returns the number of fusions made."""
    import types

    import torch.nn as nn

    from .similarity.lookup_conv import LookupConv2d, fused_motion_encoder_forward
    from .similarity.pairwise import LazyLookupPairwise4DCorr

    fused = 0
    # RAFT: lookup -> convc1 (+ bias + ReLU) as one kernel (reference models/raft.py:116-120,
    # decoder/iterative/recurrent_lookup.py:153,182,204).  The 1x1 conv is re-classed in place (same parameters and
    # state-dict keys), the encoder's forward skips the ReLU the kernel already applied, and the model's similarity class
    # hands out LazyLookup handles.
    enc = getattr(getattr(model, "update_block", None), "encoder", None)
    conv = getattr(enc, "convc1", None)
    if (
        getattr(model, "similarity_fn", None) is S.MutliScalePairwise4DCorr
        and type(conv) is nn.Conv2d and conv.kernel_size == (1, 1) and conv.stride == (1, 1) and conv.padding == (0, 0) and conv.groups == 1
        and all(hasattr(enc, a) for a in ("convf1", "convf2", "conv"))
    ):
        conv.__class__ = LookupConv2d
        enc.forward = types.MethodType(fused_motion_encoder_forward, enc)
        model.similarity_fn = LazyLookupPairwise4DCorr
        fused += 1
    for m in model.modules():
        corr = getattr(m, "correlation_layer", None)
        if not isinstance(corr, S.IterSpatialCorrelationSampler) or corr.fused_negative_slope is not None:
            continue
        for attr in ("leaky_relu", "corr_activation"):
            act = getattr(m, attr, None)
            if isinstance(act, nn.LeakyReLU) and act.negative_slope > 0:
                corr.fused_negative_slope = float(act.negative_slope)
                setattr(m, attr, nn.Identity())
                fused += 1
                break
    return fused


def uninstall():
    """Restores every binding ``install`` replaced."""
    ez_sim = importlib.import_module("ezflow.similarity")
    for (modname, attr), old in list(_saved.items()):
        if modname == "<registry>":
            if old is None:
                ez_sim.SIMILARITY_REGISTRY._obj_map.pop(attr, None)
            else:
                ez_sim.SIMILARITY_REGISTRY._obj_map[attr] = old
        elif "." in attr:  # a method of a class in that module
            cls_name, meth = attr.split(".")
            setattr(getattr(importlib.import_module(modname), cls_name), meth, old)
        elif old is not None:
            setattr(importlib.import_module(modname), attr, old)
    _saved.clear()
