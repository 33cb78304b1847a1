"""Generate tests/golden/summary_golden.json by running the REAL reference's delta-alpha summary
(vision_spectra/analysis/publication_figures.py: extract_scenario_metrics :160-259, perform_statistical_tests
:508-551) on synthetic run histories served by a fake `mlflow` module.  Build container only:

    python oracle/gen_golden_summary.py

The fixture holds the inputs (histories, accuracies) and the reference's outputs; tests/test_summary.py replays the
inputs through vision_spectra_b200.analysis.  Nothing under tests/ or the product imports the reference.
"""

from __future__ import annotations

import dataclasses
import json
import sys
import types
from pathlib import Path

import numpy as np
import pandas as pd

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, "/root/reference")


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        if len(a) == 1 and callable(a[0]) and not k and not isinstance(a[0], _Anything):
            return a[0]
        return _Anything()

    def __getattr__(self, k):
        return _Anything()


class _Stub(types.ModuleType):
    __path__: list = []

    def __getattr__(self, k):
        if k.startswith("__"):
            raise AttributeError(k)
        return _Anything()


for name in ("matplotlib", "matplotlib.pyplot", "typer", "loguru", "rich", "rich.console", "rich.table", "seaborn"):
    try:
        __import__(name)
    except Exception:
        sys.modules[name] = _Stub(name)

# ---- fake mlflow: experiments named spectral_scenario_X, a runs DataFrame, metric histories
DATA: dict[str, list[dict]] = {}


class _Metric:
    def __init__(self, step, value):
        self.step, self.value = step, value


class _Client:
    def get_metric_history(self, run_id, key):
        scen, idx = run_id.split(":")
        run = DATA[scen][int(idx)]
        hist = run["alpha"] if key == "spectral/alpha_exponent_mean" else run["stable_rank"]
        return [_Metric(s, v) for s, v in hist]


fake = types.ModuleType("mlflow")
fake.set_tracking_uri = lambda *a, **k: None
fake.MlflowClient = _Client
fake.get_experiment_by_name = lambda name: (types.SimpleNamespace(experiment_id=name.rsplit("_", 1)[1]) if name.rsplit("_", 1)[1] in DATA else None)


def _search_runs(experiment_ids, filter_string=None):
    scen = experiment_ids[0]
    rows = []
    for i, run in enumerate(DATA[scen]):
        row = {"run_id": f"{scen}:{i}"}
        if run["accuracy"] is not None:
            row["metrics.final/val_accuracy"] = run["accuracy"]
        rows.append(row)
    return pd.DataFrame(rows)


fake.search_runs = _search_runs
sys.modules["mlflow"] = fake

from vision_spectra.analysis import publication_figures as ref  # noqa: E402


def main():
    rng = np.random.default_rng(20240607)
    spec = {"A": (3, 31, 0.9, -0.35), "B": (3, 51, 0.8, -0.05), "C": (3, 51, 1.1, 0.02), "D": (4, 31, 1.0, -0.30),
            "E": (2, 31, 1.3, -0.6), "F": (1, 51, 1.2, 0.1)}
    for scen, (nruns, epochs, a0, drift) in spec.items():
        runs = []
        for r in range(nruns):
            steps = list(range(epochs))
            rng.shuffle(steps)  # histories arrive unsorted; the summary sorts by step
            keep = [s for s in steps if not (scen == "C" and r == 1 and s % 7 == 3)]  # NaN epochs were never logged
            alpha = [(int(s), float(a0 + drift * s / (epochs - 1) + 0.03 * rng.standard_normal())) for s in keep]
            sr = [(int(s), float(20.0 - 6.0 * s / (epochs - 1) + 0.2 * rng.standard_normal())) for s in steps]
            acc = None if (scen == "E" and r == 1) else float(0.7 + 0.2 * rng.random())
            runs.append({"alpha": alpha, "stable_rank": sr, "accuracy": acc})
        DATA[scen] = runs
    DATA["G"] = [{"alpha": [], "stable_rank": [(0, 5.0)], "accuracy": 0.5}]  # a run without any alpha history
    metrics = {}
    out_metrics = {}
    for scen in DATA:
        m = ref.extract_scenario_metrics(scen)
        metrics[scen] = m
        out_metrics[scen] = dataclasses.asdict(m)
    tests = ref.perform_statistical_tests(metrics)
    fixture = {"inputs": DATA, "scenario_metrics": out_metrics, "statistical_tests": tests}
    text = json.dumps(fixture, indent=1, default=lambda o: float(o)).replace("NaN", "null")
    (ROOT / "tests" / "golden" / "summary_golden.json").write_text(text)
    print("scenarios", list(out_metrics), "tests", [t["comparison"] for t in tests])


if __name__ == "__main__":
    main()
