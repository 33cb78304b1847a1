"""Drop-in for `BaseTrainer._compute_spectral_metrics` (reference training/base.py:379-416)."""

from __future__ import annotations

from typing import Any

import numpy as np
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


def save_epoch_spectral_artifacts(snapshot: Any, epoch: int, artifacts_dir: Any, mlflow_module: Any = None):
    """JSON half of `BaseTrainer._save_epoch_spectral_artifacts` (training/base.py:453-511): writes
    `spectral/json/spectral_epoch_%04d.json` with the reference's layout -- epoch, timestamp, aggregated_metrics and per
    distribution name / matrix_type / singular_values (already truncated to the tracker's max_singular_values) /
    metrics -- and logs it under `spectral/json` when an mlflow module is given.  The per-layer histogram PNGs of the
    reference (:513-567) are plotting: out of scope.  Returns the path."""
    import json
    from pathlib import Path

    json_dir = Path(artifacts_dir) / "spectral" / "json"
    json_dir.mkdir(parents=True, exist_ok=True)
    epoch_data = {
        "epoch": epoch,
        "timestamp": snapshot.timestamp,
        "aggregated_metrics": snapshot.aggregated_metrics,
        "distributions": [
            {"name": d.name, "matrix_type": d.matrix_type, "singular_values": np.asarray(d.singular_values).tolist(), "metrics": d.metrics}
            for d in snapshot.distributions
        ],
    }
    json_path = json_dir / f"spectral_epoch_{epoch:04d}.json"
    with open(json_path, "w") as f:
        json.dump(epoch_data, f, indent=2)
    if mlflow_module is not None:
        mlflow_module.log_artifact(str(json_path), artifact_path="spectral/json")
    return json_path
