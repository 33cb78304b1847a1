"""Comparison helpers with the gates of BASELINE.json / SURVEY 8d written down once."""

import math

import numpy as np

METRIC_RTOL = 1e-4  # "alpha, entropy and stable rank within 1e-4 relative"
METRIC_ABS_FLOOR = 1e-3  # absolute 1e-4 * 1e-3 when |ref| < 1e-3 (alpha ~ 0 for flat spectra)
SV_RTOL = 1e-5  # "singular values within 1e-5 relative"


def metric_close(got: float, ref: float, rtol: float = METRIC_RTOL) -> bool:
    if math.isnan(ref) or math.isnan(got):
        return math.isnan(ref) and math.isnan(got)
    return abs(got - ref) <= rtol * max(abs(ref), METRIC_ABS_FLOOR)


def sv_errors(got: np.ndarray, ref: np.ndarray) -> tuple[float, float]:
    """(normwise max|d|/s_max, elementwise max|d|/s_i)."""
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    d = np.abs(got - ref)
    smax = max(float(ref.max(initial=0.0)), 1e-300)
    return float(d.max(initial=0.0) / smax), float((d / np.maximum(ref, 1e-300)).max(initial=0.0))
