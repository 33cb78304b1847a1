"""Drop-in for `BaseTrainer._compute_spectral_metrics` (reference training/base.py:379-416)."""

from __future__ import annotations

from typing import Any

import torch

from ..engine import analyze_matrices
from ..metrics.extraction import extract_all_weights
from ..metrics.spectral import aggregate_spectral_metrics


def compute_spectral_metrics(model: torch.nn.Module, spectral_config: Any) -> dict[str, float]:
    """Overall `{metric}_{mean,std}` plus per-matrix-type `"{type}_{metric}_{mean,std}"`
    for the layers `spectral_config` selects (`layers`, `extract_qkv`, `extract_mlp`,
    `extract_patch_embed`; reference settings.py:192-223).  One batched GPU call."""
    model.eval()
    weights = extract_all_weights(
        model,
        layer_patterns=spectral_config.layers,
        include_qkv=spectral_config.extract_qkv,
        include_mlp=spectral_config.extract_mlp,
        include_patch_embed=spectral_config.extract_patch_embed,
    )
    if not weights:
        return {}
    all_metrics, _ = analyze_matrices([w.weight for w in weights], want_sv=False)
    metrics_by_type: dict[str, list[dict]] = {}
    for w, m in zip(weights, all_metrics):
        metrics_by_type.setdefault(w.matrix_type, []).append(m)
    result = aggregate_spectral_metrics(all_metrics)
    for matrix_type, type_metrics in metrics_by_type.items():
        for k, v in aggregate_spectral_metrics(type_metrics).items():
            result[f"{matrix_type}_{k}"] = v
    return result


class SpectralTrainerMixin:
    """For trainers that keep the reference's method name: expects `self.model` and
    `self.config.spectral` like `BaseTrainer` (training/base.py:41,379)."""

    def _compute_spectral_metrics(self) -> dict[str, float]:
        return compute_spectral_metrics(self.model, self.config.spectral)
