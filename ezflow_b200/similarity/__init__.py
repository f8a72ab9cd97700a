from ..registry import SIMILARITY_REGISTRY, build_similarity
from .layer import CorrelationLayer
from .matryoshka import MatryoshkaDilatedCostVolume, MatryoshkaDilatedCostVolumeList
from .lookup_conv import LazyLookup, LookupConv2d, lookup_conv1x1
from .pairwise import LazyLookupPairwise4DCorr, MutliScalePairwise4DCorr, get_pyramid_dtype, release_gradient_pool, set_pyramid_dtype
from .volume import OnDemandPairwise4DCorr, dicl_cost_volume, vcn_corr
from .sampler import (
    IterSpatialCorrelationSampler,
    WarpedFeatures,
    get_local_corr_mode,
    iter_spatial_correlation_sample,
    lazy_warp,
    set_local_corr_mode,
)

__all__ = [
    "SIMILARITY_REGISTRY",
    "build_similarity",
    "CorrelationLayer",
    "MatryoshkaDilatedCostVolume",
    "MatryoshkaDilatedCostVolumeList",
    "MutliScalePairwise4DCorr",
    "IterSpatialCorrelationSampler",
    "iter_spatial_correlation_sample",
    "set_pyramid_dtype",
    "get_pyramid_dtype",
    "release_gradient_pool",
    "set_local_corr_mode",
    "get_local_corr_mode",
    "WarpedFeatures",
    "lazy_warp",
    "LazyLookup",
    "LookupConv2d",
    "lookup_conv1x1",
    "LazyLookupPairwise4DCorr",
    "OnDemandPairwise4DCorr",
    "dicl_cost_volume",
    "vcn_corr",
]
