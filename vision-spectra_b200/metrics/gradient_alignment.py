"""Drop-in for the matrix-level functions of `vision_spectra.metrics.gradient_alignment` (reference :28-115):
`compute_rank_reducing_gradient` (U V^T, the gradient of the nuclear norm) and `compute_gradient_alignment`.  U V^T is
the polar factor of W, evaluated on the device by Newton-Schulz iterations on the FP64 tensor cores
(`lowrank.polar_factor`) -- no singular vectors are ever formed."""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np

from ..lowrank import polar_factor


@dataclass
class GradientAlignmentResult:
    """Reference gradient_alignment.py:28-45."""

    layer_name: str
    cosine_similarity: float
    training_grad_norm: float
    rank_reducing_grad_norm: float
    angle_degrees: float
    is_aligned: bool


def compute_rank_reducing_gradient(weight, rank_target: int = 1):
    """Reference :48-70.  NumPy in -> NumPy out (float64, like `U @ Vt`); torch tensor in -> float64 tensor on the device."""
    out = polar_factor(weight)
    return out.cpu().numpy() if isinstance(weight, np.ndarray) else out


def compute_gradient_alignment(training_grad, weight) -> GradientAlignmentResult:
    """Reference :73-115."""
    rank_grad = compute_rank_reducing_gradient(weight)
    to_np = lambda x: x.detach().cpu().numpy() if hasattr(x, "detach") else np.asarray(x)
    train_flat = to_np(training_grad).flatten().astype(np.float64)
    rank_flat = to_np(rank_grad).flatten().astype(np.float64)
    train_norm, rank_norm = np.linalg.norm(train_flat), np.linalg.norm(rank_flat)
    if train_norm < 1e-10 or rank_norm < 1e-10:
        return GradientAlignmentResult("", 0.0, float(train_norm), float(rank_norm), 90.0, False)
    cos_sim = float(np.clip(np.dot(train_flat, rank_flat) / (train_norm * rank_norm), -1.0, 1.0))
    return GradientAlignmentResult("", cos_sim, float(train_norm), float(rank_norm), float(np.degrees(np.arccos(cos_sim))), cos_sim > 0)
