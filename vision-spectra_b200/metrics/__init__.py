"""Hot-path subset of `vision_spectra.metrics` (reference metrics/__init__.py:28-39,60-72)."""

from .extraction import (
    WeightInfo,
    extract_all_weights,
    extract_attention_weights,
    extract_mlp_weights,
    extract_patch_embed_weights,
    extract_qkv_weights,
    group_weights_by_layer,
    group_weights_by_type,
)
from .spectral import (
    EpochSpectralSnapshot,
    SpectralDistribution,
    SpectralTracker,
    aggregate_spectral_metrics,
    alpha_exponent,
    clauset_power_law_fit,
    get_spectral_distribution,
    get_spectral_metrics,
    power_law_alpha_hill,
    spectral_entropy,
    stable_rank,
)

__all__ = [
    "spectral_entropy",
    "stable_rank",
    "alpha_exponent",
    "power_law_alpha_hill",
    "get_spectral_metrics",
    "aggregate_spectral_metrics",
    "SpectralDistribution",
    "EpochSpectralSnapshot",
    "SpectralTracker",
    "get_spectral_distribution",
    "extract_qkv_weights",
    "extract_attention_weights",
    "extract_mlp_weights",
    "extract_patch_embed_weights",
    "extract_all_weights",
    "WeightInfo",
    "group_weights_by_layer",
    "group_weights_by_type",
    "clauset_power_law_fit",  # extra: no counterpart in the reference
]
