"""Drop-in for the matrix-level functions of `vision_spectra.metrics.tail_truncation` (reference :63-152):
`truncate_weight_matrix` and `truncate_by_energy`, same signatures and `info` dicts.  The truncated matrix is W times
the spectral projector of its kept singular subspace, evaluated on the device by Newton-Schulz iterations on the FP64
tensor cores (`lowrank.spectral_truncation`); ranks and energies come from the hot path's singular values.  The
model-level experiment drivers of that file (accuracy before / after) are training code: out of scope."""

from __future__ import annotations

import numpy as np

from ..engine import default_engine
from ..lowrank import spectral_truncation


def _spectrum(weight):
    eng = default_engine(weight.device if hasattr(weight, "is_cuda") and weight.is_cuda else None)
    _, svs, _ = eng.analyze([np.asarray(weight, dtype=np.float64) if isinstance(weight, np.ndarray) else weight.double()])
    return eng, svs[0]


def _finish(weight, eng, s, k):
    total = float(np.sum(s**2))
    out = spectral_truncation(weight, k, eng, s)
    info = {
        "original_rank": int(np.sum(s > 1e-10)),
        "truncated_rank": int(k),
        "energy_retained": float(np.sum(s[:k] ** 2) / total) if total > 0 else 1.0,
    }
    if isinstance(weight, np.ndarray):
        return out.cpu().numpy().astype(weight.dtype), info
    return out.to(weight.dtype), info


def truncate_weight_matrix(weight, retention_ratio: float = 0.9, min_rank: int = 1):
    """Reference :63-105: keep k = max(min_rank, ceil(len(s) * retention_ratio)) singular values."""
    eng, s = _spectrum(weight)
    k = min(max(min_rank, int(np.ceil(len(s) * retention_ratio))), len(s))
    return _finish(weight, eng, s, k)


def truncate_by_energy(weight, energy_threshold: float = 0.99, min_rank: int = 1):
    """Reference :108-152: smallest k whose cumulative energy reaches the threshold."""
    eng, s = _spectrum(weight)
    total = float(np.sum(s**2))
    if total <= 0:
        return weight, {"original_rank": 0, "truncated_rank": 0, "energy_retained": 1.0}
    k = int(np.searchsorted(np.cumsum(s**2) / total, energy_threshold) + 1)
    k = max(min_rank, min(k, len(s)))
    return _finish(weight, eng, s, k)
