"""ezflow_b200 — B200-native (sm_100a) correlation / cost-volume hot path of neu-vi/ezflow.

Host side: Python classes mirroring ``ezflow.similarity`` (same names, constructor arguments,
``from_config`` keys, forward signatures).  Device side: hand-written CUDA behind a C ABI
(include/ezflow_b200.h, built by ``python -m ezflow_b200.build``).  No CPU fallback.
"""

from . import _cabi
from .install import fuse_model, install, uninstall
from .registry import SIMILARITY_REGISTRY, build_similarity
from .similarity import (
    CorrelationLayer,
    IterSpatialCorrelationSampler,
    LazyLookup,
    LazyLookupPairwise4DCorr,
    LookupConv2d,
    MatryoshkaDilatedCostVolume,
    MatryoshkaDilatedCostVolumeList,
    MutliScalePairwise4DCorr,
    OnDemandPairwise4DCorr,
    WarpedFeatures,
    dicl_cost_volume,
    get_local_corr_mode,
    get_pyramid_dtype,
    iter_spatial_correlation_sample,
    lazy_warp,
    lookup_conv1x1,
    release_gradient_pool,
    set_local_corr_mode,
    set_pyramid_dtype,
    vcn_corr,
)
from .utils import convex_upsample_flow, coords_grid, warp

__version__ = "0.1.0"


def library_version():
    return _cabi.lib().ezb_version()
