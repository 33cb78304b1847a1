"""Generate tests/golden/lowrank_golden.npz by running the REAL reference's singular-vector consumers
(vision_spectra/metrics/tail_truncation.py: truncate_weight_matrix :63-105, truncate_by_energy :108-152;
vision_spectra/metrics/gradient_alignment.py: compute_rank_reducing_gradient :48-70) on seeded inputs.
Build container only:

    python oracle/gen_golden_lowrank.py

The two reference files are loaded by path (their package __init__ pulls in matplotlib / timm, which are absent here; the
files themselves only need numpy, scipy and torch).  tests/test_oracle.py holds oracle/spectral_oracle.py's restatements
to these outputs; the GPU tests compare the CUDA path with the oracle on the same inputs.
"""

from __future__ import annotations

import importlib.util
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests"))
from _inputs import build_case  # noqa: E402

CASES = ["vit:C:0:q", "vit:C:0:mlp_up", "vit:E:0:mlp_down", "randn:30x50:f64", "sgd:96x384", "powerlaw:64:0.5:f32"]


def _load(name: str):
    spec = importlib.util.spec_from_file_location(f"ref_{name}", f"/root/reference/vision_spectra/metrics/{name}.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules[spec.name] = mod
    spec.loader.exec_module(mod)
    return mod


def main() -> None:
    tt, ga = _load("tail_truncation"), _load("gradient_alignment")
    out = {}
    for name in CASES:
        w = build_case(name)
        key = name.replace(":", "_")
        t90, i90 = tt.truncate_weight_matrix(w, retention_ratio=0.9)
        t50, i50 = tt.truncate_weight_matrix(w, retention_ratio=0.5)
        te, ie = tt.truncate_by_energy(w, energy_threshold=0.95)
        out[f"{key}/trunc90"] = t90
        out[f"{key}/trunc50"] = t50
        out[f"{key}/energy95"] = te
        out[f"{key}/info"] = np.array([i90["original_rank"], i90["truncated_rank"], i90["energy_retained"], i50["truncated_rank"],
                                       i50["energy_retained"], ie["truncated_rank"], ie["energy_retained"]], dtype=np.float64)
        out[f"{key}/polar"] = ga.compute_rank_reducing_gradient(w)
    np.savez_compressed(ROOT / "tests" / "golden" / "lowrank_golden.npz", **out)
    print("wrote", ROOT / "tests" / "golden" / "lowrank_golden.npz", len(out), "arrays")


if __name__ == "__main__":
    main()
