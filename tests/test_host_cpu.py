"""CPU-only host-logic tests: registry / configurable mirrors, from_config keys, buffers and state-dict
names, refusal of CPU tensors, and (when the reference checkout is present) the install() drop-in."""

import numpy as np
import pytest
import torch

import ezflow_b200 as E
from ezflow_b200.config import CfgNode


def test_registry_mirror():
    reg = E.SIMILARITY_REGISTRY
    for name in ("MutliScalePairwise4DCorr", "IterSpatialCorrelationSampler", "CorrelationLayer",
                 "MatryoshkaDilatedCostVolume", "MatryoshkaDilatedCostVolumeList"):
        assert name in reg and reg.get(name).__name__ == name
    with pytest.raises(KeyError):
        reg.get("nope")
    with pytest.raises(AssertionError):  # duplicates assert, like ezflow/utils/registry.py:23-25
        reg.register(E.CorrelationLayer)
    assert E.build_similarity(name="CorrelationLayer", instantiate=False) is E.CorrelationLayer


def test_configurable_from_config():
    cfg = CfgNode({"NAME": "IterSpatialCorrelationSampler", "KERNEL_SIZE": 1, "PATCH_SIZE": 21, "STRIDE": 1,
                   "PADDING": 0, "DILATION": 1, "DILATION_PATCH": 2})
    m = E.build_similarity(cfg)
    assert (m.patch_size, m.dilation_patch, m.kernel_size) == (21, 2, 1)
    m2 = E.IterSpatialCorrelationSampler(kernel_size=1, patch_size=9, padding=0)
    assert m2.patch_size == 9 and len(list(m2.parameters())) == 0
    c = E.CorrelationLayer(CfgNode({"PAD_SIZE": 3, "MAX_DISPLACEMENT": 3}))
    assert c.max_h_disp == 3
    assert E.MutliScalePairwise4DCorr.from_config(CfgNode({"NUM_LEVELS": 3, "CORR_RADIUS": 3})) == {
        "num_levels": 3, "corr_radius": 3}


def test_dcvnet_yaml_keys_and_buffers():
    # configs/models/dcvnet.yaml:9-16
    cfg = CfgNode({"NAME": "MatryoshkaDilatedCostVolumeList", "NUM_GROUPS": 1, "MAX_DISPLACEMENT": 4,
                   "ENCODER_OUTPUT_STRIDES": [2, 8], "DILATIONS": [[1], [1, 2, 3, 5, 9, 16]],
                   "NORMALIZE_FEAT_L2": False, "USE_RELU": False})
    m = E.build_similarity(cfg)
    assert m.get_search_range() == 9
    assert sorted(m.state_dict().keys()) == ["cost_volume_list.0.offsets", "cost_volume_list.1.offsets", "offsets_2d"]
    o2 = m.get_global_flow_offsets()
    assert tuple(o2.shape) == (7, 9, 9, 2)
    assert o2[0, 0, 0].tolist() == [-8.0, -8.0] and o2[6, 8, 0].tolist() == [512.0, -512.0]
    assert [cv.stride for cv in m.cost_volume_list] == [4, 1]
    assert len(m.cost_volume_list[1].corr_layers) == 6
    cv = E.MatryoshkaDilatedCostVolume(max_displacement=4, dilations=[1, 2, 3, 5, 9, 16])
    assert tuple(cv.get_relative_offsets().shape) == (6, 9)
    np.testing.assert_array_equal(cv.offsets[5].numpy(), np.arange(-4, 5) * 16)


def test_no_cpu_fallback():
    x = torch.rand(1, 8, 8, 8)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        E.MutliScalePairwise4DCorr(x, x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        E.IterSpatialCorrelationSampler(patch_size=3)(x, x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        E.CorrelationLayer()(x, x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        E.MatryoshkaDilatedCostVolume()(x, x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        E.warp(x, torch.zeros(1, 2, 8, 8))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        E.convex_upsample_flow(torch.zeros(1, 2, 8, 8), torch.zeros(1, 9 * 64, 8, 8), 8)
    # parameter errors come before device errors, as in the reference (sampler.py:89-94)
    with pytest.raises(NotImplementedError):
        E.iter_spatial_correlation_sample(x, x, kernel_size=3)


def test_product_does_not_import_oracle():
    import os
    import re

    root = os.path.dirname(os.path.abspath(E.__file__))
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh")):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), f"{f} imports the oracle"


def test_coords_grid_matches_reference_convention():
    g = E.coords_grid(2, 3, 4)
    assert tuple(g.shape) == (2, 2, 3, 4)
    assert g[0, 0, 1].tolist() == [0, 1, 2, 3] and g[1, 1, :, 0].tolist() == [0, 1, 2]


@pytest.mark.reference
def test_install_into_reference_and_yaml_models():
    import refshim

    if refshim.import_reference() is None:
        pytest.skip("reference checkout not present on this machine")
    from ezflow.models import build_model
    from ezflow.similarity import SIMILARITY_REGISTRY as R

    E.install()
    try:
        assert R.get("MutliScalePairwise4DCorr") is E.MutliScalePairwise4DCorr
        raft = build_model("RAFT", cfg_path=refshim.reference_config_path("raft.yaml"), custom_cfg=True)
        assert raft.similarity_fn is E.MutliScalePairwise4DCorr
        import ezflow.models.raft as ref_raft

        assert ref_raft.convex_upsample_flow is E.convex_upsample_flow
        # lookup -> convc1 fusion: the conv is re-classed in place, the state dict is untouched
        keys = list(raft.state_dict())
        assert E.fuse_model(raft) == 1 and E.fuse_model(raft) == 0
        assert isinstance(raft.update_block.encoder.convc1, E.LookupConv2d) and raft.similarity_fn is E.LazyLookupPairwise4DCorr
        assert list(raft.state_dict()) == keys
        x = torch.randn(1, 324, 6, 8)
        conv = raft.update_block.encoder.convc1
        assert torch.equal(conv(x), torch.nn.functional.conv2d(x, conv.weight, conv.bias))  # plain tensors: an ordinary Conv2d
        fnc = build_model("FlowNetC", cfg_path=refshim.reference_config_path("flownet_c.yaml"), custom_cfg=True)
        assert isinstance(fnc.correlation_layer, E.IterSpatialCorrelationSampler)
        assert (fnc.correlation_layer.patch_size, fnc.correlation_layer.dilation_patch) == (21, 2)
        pwc = build_model("PWCNet", cfg_path=refshim.reference_config_path("pwcnet.yaml"), custom_cfg=True)
        assert isinstance(pwc.decoder.correlation_layer, E.IterSpatialCorrelationSampler)
        # fusions of the sampler's neighbours: lazy warp in the PWC decoder, LeakyReLU folded into the sampler
        import ezflow.decoder.pyramid as ref_pyramid

        assert ref_pyramid.warp is E.lazy_warp
        assert E.fuse_model(pwc) == 1 and isinstance(pwc.decoder.leaky_relu, torch.nn.Identity)
        assert pwc.decoder.correlation_layer.fused_negative_slope == pytest.approx(0.1)
        assert E.fuse_model(fnc) == 1 and isinstance(fnc.corr_activation, torch.nn.Identity)
        # VCN's inline correlation and DICL's volume assembly are rebound at class level
        import ezflow.models.vcn as ref_vcn
        import ezflow.similarity.learnable_cost as ref_lc
        from ezflow_b200.similarity import volume as VOL

        assert ref_lc.LearnableMatchingCost.forward is VOL.learnable_matching_cost_forward
        assert ref_vcn.VCN._corr_fn.__module__ == "ezflow_b200.install"
        dcv = build_model("DCVNet", cfg_path=refshim.reference_config_path("dcvnet.yaml"), custom_cfg=True)
        assert isinstance(dcv.cost_volume_list, E.MatryoshkaDilatedCostVolumeList)
        ref_keys = {k for k in dcv.state_dict() if "offsets" in k}
        assert ref_keys == {"cost_volume_list.offsets_2d", "cost_volume_list.cost_volume_list.0.offsets",
                            "cost_volume_list.cost_volume_list.1.offsets"}
    finally:
        E.uninstall()
    assert R.get("MutliScalePairwise4DCorr").__module__.startswith("ezflow.")
    import ezflow.models.vcn as ref_vcn
    import ezflow.similarity.learnable_cost as ref_lc

    assert ref_vcn.VCN._corr_fn.__module__ == "ezflow.models.vcn" and ref_lc.LearnableMatchingCost.forward.__module__.startswith("ezflow.")
