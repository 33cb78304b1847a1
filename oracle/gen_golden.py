"""Generate tests/golden/* by running the REAL reference from /root/reference.

Run in the build container only (the GPU box has no /root/reference):

    python oracle/gen_golden.py

The outputs are committed; this script is committed so they can be regenerated
and audited.  Nothing under tests/ or the product imports the reference.

What is recorded, per seeded input of tests/_inputs.py:
  * the reference's four metrics  (vision_spectra.metrics.spectral.get_spectral_metrics)
  * the reference's singular values (scipy.linalg.svd on the f64 cast, exactly as
    run_spectral_analysis.py:331-333 does)
  * the integers that the reference's estimators derive (m, [start,end), k), recomputed
    here from the reference's SVs with the formulas at spectral.py:243-256,344-353
and, per stub-ViT model: extract_and_analyze_weights (run_spectral_analysis.py:297),
BaseTrainer._compute_spectral_metrics (training/base.py:379), SpectralTracker.record_epoch
(spectral.py:647) and get_spectral_distribution (spectral.py:495).
"""

from __future__ import annotations

import json
import sys
import types
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT / "tests"))
sys.path.insert(0, "/root/reference")

# The reference's driver/trainer modules import packages that are not installed
# here; none of them is touched by the functions we call.
class _Anything:
    """Absorbs any use: call, attribute, decorator, context manager."""

    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k and not isinstance(a[0], _Anything):
            return a[0]  # used as a decorator: leave the function intact
        return _Anything()

    def __getattr__(self, k):
        return _Anything()

    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False

    def __iter__(self):
        return iter(())


class _StubModule(types.ModuleType):
    __path__: list = []

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Anything()


for _name in (
    "matplotlib",
    "matplotlib.pyplot",
    "mlflow",
    "mlflow.tracking",
    "timm",
    "torchmetrics",
    "seaborn",
    "medmnist",
    "loguru",
    "typer",
    "rich",
    "rich.console",
    "rich.table",
    "rich.progress",
    "rich.panel",
    "tqdm",
):
    if _name not in sys.modules:
        try:
            __import__(_name)
        except Exception:
            sys.modules[_name] = _StubModule(_name)

import torch  # noqa: E402
from _inputs import build_case, checksum, golden_case_names  # noqa: E402
from _vit_stub import StubViT, WrappedViT  # noqa: E402
from scipy.linalg import svd  # noqa: E402

from vision_spectra.metrics import spectral as ref  # noqa: E402


def ints_from_sv(s):
    """spectral.py:243-256 and :344-353 applied to the reference's own SVs."""
    out = {"m": 0, "start": -1, "end": -1, "k": -1}
    if s is None:
        return out
    pos = s[np.isfinite(s) & (s > 0)]
    m = int(pos.size)
    out["m"] = m
    if m >= 8:
        start = max(1, int(0.10 * m))
        end = min(max(start + 6, int(0.60 * m)), m)
        if end - start >= 2:
            out["start"], out["end"] = start, end
    lam = s**2
    n = int(np.count_nonzero(np.isfinite(lam) & (lam > 0)))
    if n >= 8:
        k = max(5, int(0.10 * n))
        out["k"] = min(k, max(5, n - 1))
    return out


def main():
    out_dir = ROOT / "tests" / "golden"
    out_dir.mkdir(parents=True, exist_ok=True)
    records, svs = {}, {}
    for name in golden_case_names():
        w = build_case(name)
        metrics = ref.get_spectral_metrics(w)
        try:
            s = svd(np.asarray(w, dtype=np.float64), compute_uv=False) if w.ndim == 2 else None
        except Exception:
            s = None
        records[name] = {
            "shape": list(w.shape),
            "dtype": str(w.dtype),
            "input_crc32": checksum(w),
            "metrics": metrics,
            "ints": ints_from_sv(s),
            "sv_len": -1 if s is None else int(s.size),
        }
        if s is not None:
            svs[name] = s
        # optional-argument variants (spectral.py:259-262, :351)
        if w.ndim == 2 and min(w.shape) >= 20 and s is not None:
            records[name]["alpha_fit_range_2_12"] = ref.alpha_exponent(np.asarray(w, np.float64), fit_range=(2, 12))
            records[name]["hill_k7"] = ref.power_law_alpha_hill(np.asarray(w, np.float64), k=7)
        print(name, metrics)
    with open(out_dir / "metrics_golden.json", "w") as f:
        json.dump(records, f, indent=1)
    np.savez_compressed(out_dir / "sv_golden.npz", **svs)

    # ---------------------------------------------------------------- model level
    from vision_spectra.experiments import run_spectral_analysis as rsa
    from vision_spectra.training import base as tb

    models = {}
    for tag, (d, depth, seed, cls, kw) in {
        "E_seed42": (32, 1, 42, StubViT, {}),
        "C_seed142": (96, 3, 142, StubViT, {}),
        "E_wrapped_seed7": (32, 2, 7, WrappedViT, {}),
        "E_sepqkv_seed3": (32, 1, 3, StubViT, {"separate_qkv": True}),
    }.items():
        model = cls(embed_dim=d, depth=depth, seed=seed, **kw)
        sd_crc = checksum(np.concatenate([p.detach().numpy().ravel() for p in model.parameters()]))
        res = rsa.extract_and_analyze_weights(model, torch.device("cpu"))

        cfg = types.SimpleNamespace(
            spectral=types.SimpleNamespace(
                layers=["blocks.0"], extract_qkv=True, extract_mlp=True, extract_patch_embed=True
            )
        )
        fake_self = types.SimpleNamespace(model=model, config=cfg)
        trainer_metrics = tb.BaseTrainer._compute_spectral_metrics(fake_self)

        tracker = ref.SpectralTracker(
            layer_patterns=["blocks.0"], include_qkv=True, include_mlp=True, include_patch_embed=True,
            max_singular_values=20,
        )
        snap = tracker.record_epoch(model, 3)
        tdict = tracker.to_dict()
        for h in tdict["history"]:
            h["timestamp"] = ""
        dist0 = snap.distributions[0]
        models[tag] = {
            "params_crc32": sd_crc,
            "analysis": res,
            "trainer_metrics": trainer_metrics,
            "tracker": tdict,
            "dist0": {
                "name": dist0.name,
                "matrix_type": dist0.matrix_type,
                "singular_values": dist0.singular_values.tolist(),
                "eigenvalues": dist0.eigenvalues.tolist(),
                "normalized_sv": dist0.normalized_sv.tolist(),
                "cumulative_variance": dist0.cumulative_variance.tolist(),
                "metrics": dist0.metrics,
            },
        }
        print(tag, res["aggregated_metrics"])
    with open(out_dir / "model_golden.json", "w") as f:
        json.dump(models, f, indent=1)

    # aggregate on a handcrafted list incl. NaNs (spectral.py:445-460)
    lst = [
        {"spectral_entropy": 1.0, "stable_rank": 2.0, "alpha_exponent": float("nan"), "pl_alpha_hill": 3.0},
        {"spectral_entropy": 2.0, "stable_rank": 5.0, "alpha_exponent": float("nan"), "pl_alpha_hill": float("inf")},
        {"spectral_entropy": 4.0, "stable_rank": 11.0, "alpha_exponent": float("nan"), "pl_alpha_hill": 1.5},
    ]
    with open(out_dir / "aggregate_golden.json", "w") as f:
        json.dump({"input": lst, "output": ref.aggregate_spectral_metrics(lst), "empty": ref.aggregate_spectral_metrics([])}, f, indent=1)
    print("golden written to", out_dir)


if __name__ == "__main__":
    main()
