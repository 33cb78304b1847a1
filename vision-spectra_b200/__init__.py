"""vision-spectra_b200: B200-native weight-spectrum analysis path of mgrts/vision-spectra.

Importable as `vision_spectra_b200` (the repo-root alias package points here).
Only the hot path lives here: CUDA kernels + C-ABI (`csrc/`, `include/vspectra.h`),
the batched engine, and host-side mirrors of the reference's metric / extraction /
caller interfaces.
"""

from . import _native
from .engine import METRIC_KEYS, SpectraEngine, analyze_matrices, default_engine, nan_metrics

__all__ = ["SpectraEngine", "analyze_matrices", "default_engine", "nan_metrics", "METRIC_KEYS", "_native"]
__version__ = "0.1.0"
