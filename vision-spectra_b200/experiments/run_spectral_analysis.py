"""Drop-in for the hot-path part of `vision_spectra.experiments.run_spectral_analysis`:
`extract_and_analyze_weights` (reference :297-345), the MLflow metric logging the
driver does with its result (:511-513, :586-588) and the JSON artifacts of
`log_spectral_artifacts` (:348-384).  The training loop, scenario table and CLI of
that file are out of scope (SURVEY 8 scope table).
"""

from __future__ import annotations

import json
from pathlib import Path
from typing import Any

import numpy as np
import torch

from ..engine import analyze_matrices
from ..metrics.extraction import extract_attention_weights, extract_mlp_weights, extract_qkv_weights
from ..metrics.spectral import aggregate_spectral_metrics


def extract_and_analyze_weights(model: torch.nn.Module, device: torch.device | None = None) -> dict[str, Any]:
    """Same result dict as the reference (:341-345):
    `per_layer_metrics` {name: 4-key dict}, `aggregated_metrics` {key_mean/_std},
    `singular_values` {name: descending list of all min(r,c) SVs, [] if the SVD
    failed (:335-336)} -- computed by ONE batched GPU call instead of 5 CPU SVDs
    per matrix.  `device` is accepted for signature compatibility; the matrices are
    analysed on the device they live on (CPU models are uploaded once)."""
    model.eval()
    all_weights = extract_qkv_weights(model) + extract_attention_weights(model) + extract_mlp_weights(model)
    metrics, svs = analyze_matrices([w.weight for w in all_weights]) if all_weights else ([], [])
    per_layer_metrics: dict[str, dict[str, float]] = {}
    singular_values: dict[str, list[float]] = {}
    layer_metrics_list = []
    for w, m, s in zip(all_weights, metrics, svs):
        per_layer_metrics[w.name] = m
        layer_metrics_list.append(m)
        singular_values[w.name] = [] if s is None else s.tolist()
    return {
        "per_layer_metrics": per_layer_metrics,
        "aggregated_metrics": aggregate_spectral_metrics(layer_metrics_list),
        "singular_values": singular_values,
    }


def log_spectral_metrics(mlflow_module: Any, analysis: dict[str, Any], epoch: int) -> int:
    """`mlflow.log_metric(f"spectral/{key}", value, step=epoch)` for every finite
    aggregated value -- the reference's lines :511-513 / :586-588.  `mlflow_module`
    is passed in (mlflow itself, or a recorder in tests).  Returns how many metrics
    were logged."""
    n = 0
    for key, value in analysis["aggregated_metrics"].items():
        if np.isfinite(value):
            mlflow_module.log_metric(f"spectral/{key}", value, step=epoch)
            n += 1
    return n


def write_spectral_artifacts(analysis: dict[str, Any], epoch: int, out_dir: str | Path, mlflow_module: Any = None) -> Path:
    """`epoch_{N}/singular_values.json` and `epoch_{N}/layer_metrics.json` (NaN ->
    null) exactly as reference :364-384; logged under `spectral/epoch_{N}` when an
    mlflow module is given.  Histogram PNGs (:386-412) are plotting, out of scope."""
    epoch_dir = Path(out_dir) / f"epoch_{epoch}"
    epoch_dir.mkdir(parents=True, exist_ok=True)
    values_file = epoch_dir / "singular_values.json"
    with open(values_file, "w") as f:
        json.dump(analysis["singular_values"], f, indent=2)
    metrics_file = epoch_dir / "layer_metrics.json"
    clean = {
        layer: {k: v if np.isfinite(v) else None for k, v in metrics.items()}
        for layer, metrics in analysis["per_layer_metrics"].items()
    }
    with open(metrics_file, "w") as f:
        json.dump(clean, f, indent=2)
    if mlflow_module is not None:
        mlflow_module.log_artifact(str(values_file), f"spectral/epoch_{epoch}")
        mlflow_module.log_artifact(str(metrics_file), f"spectral/epoch_{epoch}")
    return epoch_dir
