"""Drop-in for `vision_spectra.metrics.spectral` (reference file of the same name).

Same function names, argument meaning, return types and failure convention (NaN /
None, never raise for bad data) as the reference -- but the singular values and
the four metrics come from the sm_100a CUDA path (`engine.SpectraEngine`) instead
of scipy.linalg.svd on the CPU.  Per-function reference lines are cited in each
docstring.  The batched call (`engine.analyze_matrices`) is what callers should
use per checkpoint; the single-matrix functions here are thin wrappers for
signature compatibility and pay one launch sequence per call.
"""

from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Any

import numpy as np

from ..engine import METRIC_KEYS, analyze_matrices, nan_metrics

__all__ = [
    "spectral_entropy",
    "stable_rank",
    "alpha_exponent",
    "power_law_alpha_hill",
    "get_spectral_metrics",
    "aggregate_spectral_metrics",
    "SpectralDistribution",
    "get_spectral_distribution",
    "EpochSpectralSnapshot",
    "SpectralTracker",
    "distribution_from_sv",
]

_NAN = float("nan")


def _is_2d(w: Any) -> bool:
    return getattr(w, "ndim", None) == 2


def _one(w: Any, key: str, **kw) -> float:
    if not _is_2d(w):
        return _NAN
    metrics, _ = analyze_matrices([w], want_sv=False, **kw)
    return metrics[0][key]


def spectral_entropy(weight_matrix) -> float:
    """Shannon entropy (nats) of s_i^2 / sum s^2.  Reference spectral.py:49-109."""
    return _one(weight_matrix, "spectral_entropy")


def stable_rank(weight_matrix) -> float:
    """sum s^2 / max(s)^2.  Reference spectral.py:112-173."""
    return _one(weight_matrix, "stable_rank")


def alpha_exponent(weight_matrix, fit_range: tuple[int, int] | None = None) -> float:
    """-slope of ln(s_i) on ln(rank) over the index window [start, end); default
    window [10 %, 60 %) of the m positive SVs, NaN if m < 8.  Reference
    spectral.py:176-273."""
    if fit_range is not None and (fit_range[0] < 0 or fit_range[1] < 0):
        return _NAN  # the reference's negative-index slicing is not reproduced
    return _one(weight_matrix, "alpha_exponent", fit_range=fit_range)


def power_law_alpha_hill(weight_matrix, k: int | None = None) -> float:
    """Hill tail index 1 + 1/mean(ln(lambda_(i)/lambda_(k))) on the k largest
    eigenvalues lambda = s^2; default k = top 10 % (min 5).  Reference
    spectral.py:276-368."""
    if k is not None and k < 1:
        return _NAN
    return _one(weight_matrix, "pl_alpha_hill", hill_k=k)


def get_spectral_metrics(weight_matrix) -> dict[str, float]:
    """All four metrics, keys in the reference's order.  Reference spectral.py:371-414
    (accepts a NumPy array or anything tensor-like; non-2-D input -> all NaN)."""
    if not _is_2d(weight_matrix):
        return nan_metrics()
    metrics, _ = analyze_matrices([weight_matrix], want_sv=False)
    return metrics[0]


def aggregate_spectral_metrics(metrics_list: list[dict[str, float]]) -> dict[str, float]:
    """Mean and population std over the finite entries of each key; key set and
    order from the first dict; `{}` for an empty list.  Reference spectral.py:417-460.
    (<= 72 scalars per checkpoint: host arithmetic, as in the reference.)"""
    if not metrics_list:
        return {}
    result: dict[str, float] = {}
    for key in metrics_list[0]:
        values = [m[key] for m in metrics_list if np.isfinite(m.get(key, np.nan))]
        if values:
            result[f"{key}_mean"] = float(np.mean(values))
            result[f"{key}_std"] = float(np.std(values))
        else:
            result[f"{key}_mean"] = np.nan
            result[f"{key}_std"] = np.nan
    return result


@dataclass
class SpectralDistribution:
    """Same fields as the reference container, spectral.py:468-492."""

    name: str
    matrix_type: str
    singular_values: np.ndarray
    eigenvalues: np.ndarray
    normalized_sv: np.ndarray
    cumulative_variance: np.ndarray
    metrics: dict[str, float]


def distribution_from_sv(
    s: np.ndarray | None, metrics: dict[str, float], name: str = "", matrix_type: str = "unknown"
) -> SpectralDistribution | None:
    """Derived arrays of spectral.py:540-570 from device-computed singular values."""
    if s is None:
        return None
    s = np.asarray(s, dtype=np.float64)
    s = s[np.isfinite(s) & (s >= 0)]
    if s.size == 0:
        return None
    s = np.sort(s)[::-1]
    eigenvalues = s**2
    s_max = s[0] if s[0] > 0 else 1.0
    total = eigenvalues.sum()
    cumvar = np.cumsum(eigenvalues) / total if total > 0 else np.zeros_like(eigenvalues)
    return SpectralDistribution(
        name=name,
        matrix_type=matrix_type,
        singular_values=s,
        eigenvalues=eigenvalues,
        normalized_sv=s / s_max,
        cumulative_variance=cumvar,
        metrics=metrics,
    )


def distribution_from_device(d: np.ndarray | None, metrics: dict[str, float], name: str = "", matrix_type: str = "unknown") -> SpectralDistribution | None:
    """SpectralDistribution from the [4, k] block the metrics kernel wrote (rows: singular values, eigenvalues,
    normalized_sv, cumulative_variance -- spectral.py:545-557 computed on the device, truncated to k entries)."""
    if d is None:
        return None
    return SpectralDistribution(name=name, matrix_type=matrix_type, singular_values=d[0], eigenvalues=d[1],
                                normalized_sv=d[2], cumulative_variance=d[3], metrics=metrics)


def get_spectral_distribution(weight_matrix, name: str = "", matrix_type: str = "unknown") -> SpectralDistribution | None:
    """Full distribution of one matrix; None for non-2-D input or failed SVD.
    Reference spectral.py:495-570.  The four arrays come from the device (full length: dist_k = min(rows, cols))."""
    if not _is_2d(weight_matrix):
        return None
    from ..engine import default_engine

    dev = weight_matrix.device if hasattr(weight_matrix, "device") and getattr(weight_matrix.device, "type", "") == "cuda" else None
    dists: list = []
    metrics, _, _ = default_engine(dev).analyze([weight_matrix], want_sv=False, dist_k=max(1, min(weight_matrix.shape)), dist_out=dists)
    return distribution_from_device(dists[0], metrics[0], name, matrix_type)


@dataclass
class EpochSpectralSnapshot:
    """Reference spectral.py:573-594."""

    epoch: int
    distributions: list[SpectralDistribution]
    aggregated_metrics: dict[str, float]
    timestamp: str = ""

    def __post_init__(self):
        if not self.timestamp:
            from datetime import datetime

            self.timestamp = datetime.now().isoformat()


class SpectralTracker:
    """Per-epoch history of spectral distributions; same constructor arguments,
    methods and JSON layout as the reference (spectral.py:597-843), but
    `record_epoch` sends every selected matrix to the GPU in ONE batch."""

    def __init__(
        self,
        layer_patterns: list[str] | None = None,
        include_qkv: bool = True,
        include_mlp: bool = False,
        include_patch_embed: bool = True,
        max_singular_values: int = 100,
    ):
        self.layer_patterns = layer_patterns or []
        self.include_qkv = include_qkv
        self.include_mlp = include_mlp
        self.include_patch_embed = include_patch_embed
        self.max_singular_values = max_singular_values
        self.history: list[EpochSpectralSnapshot] = []

    def record_epoch(self, model: Any, epoch: int) -> EpochSpectralSnapshot:
        """Reference spectral.py:647-706."""
        from .extraction import extract_all_weights

        weights = extract_all_weights(
            model,
            layer_patterns=self.layer_patterns,
            include_qkv=self.include_qkv,
            include_mlp=self.include_mlp,
            include_patch_embed=self.include_patch_embed,
        )
        # one batched call; the truncation to max_singular_values (reference :683-692) happens in the metrics kernel,
        # so 4 k values per matrix come back instead of min(rows, cols) singular values
        from ..engine import default_engine

        distributions = []
        if weights:
            dev = next((w.weight.device for w in weights if getattr(w.weight, "is_cuda", False)), None)
            dists: list = []
            metrics, _, _ = default_engine(dev).analyze([w.weight for w in weights], want_sv=False,
                                                        dist_k=max(1, int(self.max_singular_values)), dist_out=dists)
            for w, m, d in zip(weights, metrics, dists):
                dist = distribution_from_device(d, m, w.name, w.matrix_type)
                if dist is not None:
                    distributions.append(dist)
        all_metrics = [d.metrics for d in distributions]
        snapshot = EpochSpectralSnapshot(
            epoch=epoch,
            distributions=distributions,
            aggregated_metrics=aggregate_spectral_metrics(all_metrics) if all_metrics else {},
        )
        self.history.append(snapshot)
        return snapshot

    def get_metric_history(self, metric_name: str) -> tuple[list[int], list[float]]:
        """Reference spectral.py:708-728."""
        epochs, values = [], []
        for snap in self.history:
            if metric_name in snap.aggregated_metrics:
                v = snap.aggregated_metrics[metric_name]
                if np.isfinite(v):
                    epochs.append(snap.epoch)
                    values.append(v)
        return epochs, values

    def get_layer_sv_history(self, layer_name: str) -> tuple[list[int], list[np.ndarray]]:
        """Reference spectral.py:730-750."""
        epochs, out = [], []
        for snap in self.history:
            for dist in snap.distributions:
                if dist.name == layer_name:
                    epochs.append(snap.epoch)
                    out.append(dist.singular_values)
                    break
        return epochs, out

    def get_all_layer_names(self) -> list[str]:
        """Reference spectral.py:752-756."""
        return [d.name for d in self.history[0].distributions] if self.history else []

    def to_dict(self) -> dict[str, Any]:
        """Reference spectral.py:758-788 (same keys, same nesting)."""
        return {
            "layer_patterns": self.layer_patterns,
            "include_qkv": self.include_qkv,
            "include_mlp": self.include_mlp,
            "include_patch_embed": self.include_patch_embed,
            "max_singular_values": self.max_singular_values,
            "history": [
                {
                    "epoch": s.epoch,
                    "timestamp": s.timestamp,
                    "aggregated_metrics": s.aggregated_metrics,
                    "distributions": [
                        {
                            "name": d.name,
                            "matrix_type": d.matrix_type,
                            "singular_values": d.singular_values.tolist(),
                            "metrics": d.metrics,
                        }
                        for d in s.distributions
                    ],
                }
                for s in self.history
            ],
        }

    def save(self, path: Path) -> None:
        """Reference spectral.py:790-798."""
        import json

        path = Path(path)
        path.parent.mkdir(parents=True, exist_ok=True)
        with open(path, "w") as f:
            json.dump(self.to_dict(), f, indent=2)

    @classmethod
    def load(cls, path: Path) -> "SpectralTracker":
        """Reference spectral.py:800-843, including its (sum s)^2 normalisation of the
        reloaded cumulative variance."""
        import json

        with open(path) as f:
            data = json.load(f)
        tracker = cls(
            layer_patterns=data.get("layer_patterns", []),
            include_qkv=data.get("include_qkv", True),
            include_mlp=data.get("include_mlp", False),
            include_patch_embed=data.get("include_patch_embed", True),
            max_singular_values=data.get("max_singular_values", 100),
        )
        for h in data.get("history", []):
            dists = []
            for d in h.get("distributions", []):
                sv = np.array(d["singular_values"])
                dists.append(
                    SpectralDistribution(
                        name=d["name"],
                        matrix_type=d["matrix_type"],
                        singular_values=sv,
                        eigenvalues=sv**2,
                        normalized_sv=sv / sv[0] if sv[0] > 0 else sv,
                        cumulative_variance=np.cumsum(sv**2) / sv.sum() ** 2 if sv.sum() > 0 else np.zeros_like(sv),
                        metrics=d.get("metrics", {}),
                    )
                )
            tracker.history.append(
                EpochSpectralSnapshot(
                    epoch=h["epoch"],
                    distributions=dists,
                    aggregated_metrics=h.get("aggregated_metrics", {}),
                    timestamp=h.get("timestamp", ""),
                )
            )
        return tracker


def clauset_power_law_fit(weight_matrix) -> dict[str, float]:
    """Extra output (BASELINE north_star stage 3; the reference has no counterpart): Clauset-Shalizi-Newman
    x_min scan on the eigenvalue spectrum -- MLE alpha and Kolmogorov-Smirnov distance for every candidate cutoff in
    parallel on the device, the best cutoff returned: alpha, xmin, ks_distance, xmin_index (0 = largest eigenvalue),
    tail_count.  NaN / -1 for non-2-D input or fewer than eight positive eigenvalues."""
    nan = float("nan")
    if not _is_2d(weight_matrix):
        return {"alpha": nan, "xmin": nan, "ks_distance": nan, "xmin_index": -1, "tail_count": -1}
    from ..engine import default_engine

    dev = weight_matrix.device if getattr(getattr(weight_matrix, "device", None), "type", "") == "cuda" else None
    out: list = []
    default_engine(dev).analyze([weight_matrix], want_sv=False, clauset_out=out)
    return out[0] if out[0] is not None else {"alpha": nan, "xmin": nan, "ks_distance": nan, "xmin_index": -1, "tail_count": -1}
