"""Per-scenario delta-alpha summary and the pairwise significance tests of the six-scenario report, computed
from the per-epoch aggregated metrics this package already holds -- no MLflow round trip.

Reference semantics restated (vision_spectra/analysis/publication_figures.py):
  * ScenarioMetrics fields and SCENARIO_METADATA                                  :109-135
  * per run: history of `spectral/alpha_exponent_mean` sorted by step; initial = first entry, final = last
    entry, delta = final - initial; same for `spectral/stable_rank_mean`          :199-236
    (only finite values were ever logged: run_spectral_analysis.py:511-513, so NaN epochs are absent)
  * aggregates: np.mean of the per-run values, delta_alpha_std = np.std (population, ddof = 0);
    accuracy mean / std come from a pandas column, i.e. NaN-skipping mean and *sample* std (ddof = 1) :188-196,239-245
  * tests: scipy.stats.ttest_ind (Student, pooled variance, two-sided) on the delta-alpha lists of the pairs
    A-B, D-C, E-F, B-C, C-F, A-F; skipped when a list has fewer than two values; mean_diff = mean(s2) - mean(s1),
    significant = p < 0.05, and the three interpretation strings                   :508-551
SciPy is not a dependency of the product: the t distribution's tail comes from the regularised incomplete beta
function evaluated here (Lentz continued fraction).
"""

from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Iterable, Mapping, Sequence

import numpy as np


@dataclass
class ScenarioMetrics:
    """Aggregated metrics for a scenario (same fields as the reference's dataclass)."""

    scenario: str
    name: str
    description: str
    accuracy_mean: float
    accuracy_std: float
    alpha_initial_mean: float
    alpha_final_mean: float
    delta_alpha_mean: float
    delta_alpha_std: float
    delta_alpha_values: list[float]
    stable_rank_initial_mean: float
    stable_rank_final_mean: float
    num_runs: int


SCENARIO_METADATA: dict[str, dict[str, str]] = {
    "A": {"name": "Expressive+Simple", "description": "Large network on simple synthetic data"},
    "B": {"name": "Expressive+Complex", "description": "Large network on complex PathMNIST data"},
    "C": {"name": "Reduced+Complex", "description": "Reduced network on complex data"},
    "D": {"name": "Reduced+Simple", "description": "Reduced network on simple data"},
    "E": {"name": "Tiny+Simple", "description": "Minimal network on simple data"},
    "F": {"name": "Tiny+Complex", "description": "Minimal network on complex data"},
}

TEST_PAIRS = [("A", "B"), ("D", "C"), ("E", "F"), ("B", "C"), ("C", "F"), ("A", "F")]


@dataclass
class RunHistory:
    """What one finished run contributes: (step, value) histories of the two aggregated metrics and its final
    validation accuracy (None / NaN if it was not logged)."""

    alpha: list[tuple[int, float]] = field(default_factory=list)
    stable_rank: list[tuple[int, float]] = field(default_factory=list)
    accuracy: float | None = None


def run_history_from_epochs(aggregated_by_epoch: Mapping[int, Mapping[str, float]], accuracy: float | None = None) -> RunHistory:
    """Build a RunHistory from {epoch: aggregate_spectral_metrics(...)} as produced per checkpoint by this
    package; non-finite values are dropped, exactly as they would never have been logged."""
    h = RunHistory(accuracy=accuracy)
    for epoch in aggregated_by_epoch:
        agg = aggregated_by_epoch[epoch]
        a, s = agg.get("alpha_exponent_mean"), agg.get("stable_rank_mean")
        if a is not None and np.isfinite(a):
            h.alpha.append((int(epoch), float(a)))
        if s is not None and np.isfinite(s):
            h.stable_rank.append((int(epoch), float(s)))
    return h


def _first_last(history: Sequence[tuple[int, float]]) -> tuple[float, float] | None:
    if not history:
        return None
    ordered = sorted(history, key=lambda sv: sv[0])  # stable, like the reference's sort by step
    return ordered[0][1], ordered[-1][1]


def scenario_metrics(scenario: str, runs: Iterable[RunHistory]) -> ScenarioMetrics | None:
    """ScenarioMetrics of one scenario from its finished runs; None if there are none."""
    runs = list(runs)
    if not runs:
        return None
    acc = np.array([np.nan if r.accuracy is None else float(r.accuracy) for r in runs], dtype=np.float64)
    have = acc[np.isfinite(acc)]
    accuracy_mean = float(have.mean()) if have.size else float("nan")
    accuracy_std = float(have.std(ddof=1)) if have.size > 1 else float("nan")  # pandas Series.std
    a_init, a_final, deltas, s_init, s_final = [], [], [], [], []
    for r in runs:
        fl = _first_last(r.alpha)
        if fl is not None:
            a_init.append(fl[0])
            a_final.append(fl[1])
            deltas.append(fl[1] - fl[0])
        fl = _first_last(r.stable_rank)
        if fl is not None:
            s_init.append(fl[0])
            s_final.append(fl[1])
    mean = lambda v: float(np.mean(v)) if v else float("nan")  # noqa: E731
    meta = SCENARIO_METADATA.get(scenario, {"name": scenario, "description": ""})
    return ScenarioMetrics(
        scenario=scenario,
        name=meta["name"],
        description=meta["description"],
        accuracy_mean=accuracy_mean,
        accuracy_std=accuracy_std,
        alpha_initial_mean=mean(a_init),
        alpha_final_mean=mean(a_final),
        delta_alpha_mean=mean(deltas),
        delta_alpha_std=float(np.std(deltas)) if deltas else float("nan"),
        delta_alpha_values=deltas,
        stable_rank_initial_mean=mean(s_init),
        stable_rank_final_mean=mean(s_final),
        num_runs=len(runs),
    )


# ------------------------------------------------------------------------------------- Student's t
def _betacf(a: float, b: float, x: float) -> float:
    """Continued fraction of the incomplete beta function (modified Lentz)."""
    tiny = 1e-300
    qab, qap, qam = a + b, a + 1.0, a - 1.0
    c, d = 1.0, 1.0 - qab * x / qap
    d = tiny if abs(d) < tiny else d
    d = 1.0 / d
    h = d
    for m in range(1, 10000):
        m2 = 2 * m
        aa = m * (b - m) * x / ((qam + m2) * (a + m2))
        d = 1.0 + aa * d
        d = tiny if abs(d) < tiny else d
        c = 1.0 + aa / c
        c = tiny if abs(c) < tiny else c
        d = 1.0 / d
        h *= d * c
        aa = -(a + m) * (qab + m) * x / ((a + m2) * (qap + m2))
        d = 1.0 + aa * d
        d = tiny if abs(d) < tiny else d
        c = 1.0 + aa / c
        c = tiny if abs(c) < tiny else c
        d = 1.0 / d
        delta = d * c
        h *= delta
        if abs(delta - 1.0) < 1e-16:
            break
    return h


def _betainc(a: float, b: float, x: float) -> float:
    """Regularised incomplete beta function I_x(a, b)."""
    if x <= 0.0:
        return 0.0
    if x >= 1.0:
        return 1.0
    ln_front = math.lgamma(a + b) - math.lgamma(a) - math.lgamma(b) + a * math.log(x) + b * math.log1p(-x)
    if x < (a + 1.0) / (a + b + 2.0):
        return math.exp(ln_front) * _betacf(a, b, x) / a
    return 1.0 - math.exp(ln_front) * _betacf(b, a, 1.0 - x) / b


def students_t_test(x: Sequence[float], y: Sequence[float]) -> tuple[float, float]:
    """Two-sided independent two-sample t-test with pooled variance (scipy.stats.ttest_ind's default)."""
    x, y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
    n1, n2 = x.size, y.size
    df = n1 + n2 - 2
    v1, v2 = x.var(ddof=1), y.var(ddof=1)
    sp2 = ((n1 - 1) * v1 + (n2 - 1) * v2) / df
    denom = math.sqrt(sp2 * (1.0 / n1 + 1.0 / n2))
    diff = float(x.mean() - y.mean())
    if denom == 0.0:
        t = float("nan") if diff == 0.0 else math.copysign(float("inf"), diff)
        return t, (float("nan") if diff == 0.0 else 0.0)
    t = diff / denom
    p = _betainc(0.5 * df, 0.5, df / (df + t * t))
    return float(t), float(p)


def perform_statistical_tests(metrics: Mapping[str, ScenarioMetrics]) -> list[dict]:
    """Pairwise tests between scenarios on their delta-alpha values (same pairs, keys and strings as the reference)."""
    results = []
    for s1, s2 in TEST_PAIRS:
        if s1 not in metrics or s2 not in metrics:
            continue
        vals1, vals2 = metrics[s1].delta_alpha_values, metrics[s2].delta_alpha_values
        if len(vals1) < 2 or len(vals2) < 2:
            continue
        t_stat, p_value = students_t_test(vals1, vals2)
        diff = float(np.mean(vals2) - np.mean(vals1))
        significant = bool(p_value < 0.05)
        interpretation = "No significant difference"
        if significant and diff > 0:
            interpretation = f"{s2} has significantly higher compression"
        elif significant and diff < 0:
            interpretation = f"{s1} has significantly higher compression"
        results.append(
            {
                "comparison": f"{s1} vs {s2}",
                "mean_diff": diff,
                "t_statistic": float(t_stat),
                "p_value": float(p_value),
                "significant": significant,
                "interpretation": interpretation,
            }
        )
    return results
