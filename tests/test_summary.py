"""vision_spectra_b200.analysis (delta-alpha summary) against the reference's own
extract_scenario_metrics / perform_statistical_tests outputs (tests/golden/summary_golden.json, made by
oracle/gen_golden_summary.py from the real reference behind a fake MLflow)."""

import json
import math
from pathlib import Path

import numpy as np
import pytest

from vision_spectra_b200.analysis import (
    RunHistory,
    perform_statistical_tests,
    run_history_from_epochs,
    scenario_metrics,
    students_t_test,
)

GOLD = json.loads((Path(__file__).parent / "golden" / "summary_golden.json").read_text())


def _close(a, b, tol=1e-12):
    if b is None or (isinstance(b, float) and math.isnan(b)):
        return a is None or math.isnan(a)
    return abs(a - b) <= tol * max(1.0, abs(b))


def _metrics():
    out = {}
    for scen, runs in GOLD["inputs"].items():
        hist = [RunHistory([tuple(x) for x in r["alpha"]], [tuple(x) for x in r["stable_rank"]], r["accuracy"]) for r in runs]
        out[scen] = scenario_metrics(scen, hist)
    return out


def test_scenario_metrics_match_reference():
    got = _metrics()
    assert list(got) == list(GOLD["scenario_metrics"])
    for scen, ref in GOLD["scenario_metrics"].items():
        m = got[scen]
        for key, val in ref.items():
            mine = getattr(m, key)
            if key == "delta_alpha_values":
                assert len(mine) == len(val) and all(_close(a, b) for a, b in zip(mine, val)), (scen, key)
            elif isinstance(val, (int, str)) and not isinstance(val, bool) and key in ("scenario", "name", "description", "num_runs"):
                assert mine == val, (scen, key)
            else:
                assert _close(mine, val), (scen, key, mine, val)


def test_statistical_tests_match_reference():
    got = perform_statistical_tests(_metrics())
    ref = GOLD["statistical_tests"]
    assert [g["comparison"] for g in got] == [r["comparison"] for r in ref]
    for g, r in zip(got, ref):
        assert g["significant"] == r["significant"] and g["interpretation"] == r["interpretation"]
        assert _close(g["mean_diff"], r["mean_diff"]) and _close(g["t_statistic"], r["t_statistic"], 1e-11)
        assert abs(g["p_value"] - r["p_value"]) <= 1e-10 * max(r["p_value"], 1e-3)


def test_t_test_known_values_and_edges():
    # textbook example: equal means -> t = 0, p = 1
    t, p = students_t_test([1.0, 2.0, 3.0], [1.0, 2.0, 3.0])
    assert t == 0.0 and abs(p - 1.0) < 1e-15
    # identical constant samples: undefined; different constants: infinitely significant
    t, p = students_t_test([1.0, 1.0], [1.0, 1.0])
    assert math.isnan(t) and math.isnan(p)
    t, p = students_t_test([2.0, 2.0], [1.0, 1.0])
    assert math.isinf(t) and p == 0.0


def test_history_from_epochs_drops_non_finite():
    h = run_history_from_epochs({2: {"alpha_exponent_mean": 0.5, "stable_rank_mean": 3.0}, 0: {"alpha_exponent_mean": float("nan"), "stable_rank_mean": 4.0},
                                 1: {"alpha_exponent_mean": 0.7}}, accuracy=0.9)
    assert sorted(h.alpha) == [(1, 0.7), (2, 0.5)] and sorted(h.stable_rank) == [(0, 4.0), (2, 3.0)]
    m = scenario_metrics("A", [h])
    assert m.num_runs == 1 and abs(m.delta_alpha_mean - (0.5 - 0.7)) < 1e-15 and math.isnan(m.accuracy_std)
    assert scenario_metrics("A", []) is None
