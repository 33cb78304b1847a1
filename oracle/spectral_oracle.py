"""CPU oracle for the weight-spectrum analysis path.  TEST INFRASTRUCTURE ONLY.

This module restates, in NumPy/SciPy, the algorithm of the reference's hot path
(`vision_spectra/metrics/spectral.py`, `vision_spectra/metrics/extraction.py`,
`vision_spectra/experiments/run_spectral_analysis.py:297-345`).  It exists so
that the CUDA path can be checked on the GPU box, where `/root/reference` does
not exist.  Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s
`cpu_baseline` / `--impl reference` legs may import it.  The product package
(`vision-spectra_b200/`) never imports anything from `oracle/`.

Parity pinning: `oracle/gen_golden.py` imports the *real* reference from
`/root/reference` in the build container, runs it on seeded inputs and commits
the outputs under `tests/golden/`; `tests/test_oracle.py` checks this
restatement against those fixtures (bit-level for integers, 1e-12 for floats)
and against the reference's own known-answer tests
(`/root/reference/tests/test_metrics.py:12-220,279-353`).

Third-party arithmetic (not under `/root/reference`): `scipy.linalg.svd`
(LAPACK dgesdd, float64; SciPy pinned 1.17.0 in `poetry.lock:3749`, 1.18.1 in
this image), `scipy.stats.entropy`, `numpy.polyfit`.  The oracle calls the same
library entry points as the reference.

Differences from the reference on purpose: one SVD per matrix is shared by all
four metrics (the reference recomputes it 4-5 times, `spectral.py:91,158,239,339`
+ `run_spectral_analysis.py:333`; same input -> identical output, so the result
is unchanged).  `reference_cost_svds` says how many SVDs the reference would
have paid, for the cpu_baseline timing which must charge them.
"""

from __future__ import annotations

import re
from dataclasses import dataclass

import numpy as np
from scipy.linalg import svd as _svd
from scipy.stats import entropy as _scipy_entropy

METRIC_KEYS = ("spectral_entropy", "stable_rank", "alpha_exponent", "pl_alpha_hill")


# --------------------------------------------------------------------------- SVD
def singular_values(w) -> np.ndarray | None:
    """f64 singular values, descending, or None on failure.

    Follows spectral.py:91 (`svd(weight_matrix, compute_uv=False)` inside
    try/except) on the f64 cast of spectral.py:405-407.
    """
    w = np.asarray(w, dtype=np.float64)
    if w.ndim != 2:
        return None
    try:
        return _svd(w, compute_uv=False)
    except Exception:
        return None


# ----------------------------------------------------------------- scalar metrics
def entropy_from_sv(s: np.ndarray) -> float:
    """spectral.py:95-109."""
    s = s[np.isfinite(s) & (s > 0)]
    if s.size == 0:
        return float("nan")
    p = (s**2).astype(np.float64)
    total = p.sum()
    if total <= 0 or not np.isfinite(total):
        return float("nan")
    return float(_scipy_entropy(p / total))


def stable_rank_from_sv(s: np.ndarray) -> float:
    """spectral.py:162-173."""
    s = s[np.isfinite(s) & (s >= 0)]
    if s.size == 0:
        return float("nan")
    s_max = s.max()
    if s_max <= 0 or not np.isfinite(s_max):
        return float("nan")
    return float(np.sum(s**2)) / float(s_max**2)


def alpha_window(m: int, fit_range=None):
    """(start, end) of the OLS window or None -> NaN.  spectral.py:247-262."""
    if m == 0:
        return None
    if fit_range is None:
        if m < 8:
            return None
        start = max(1, int(0.10 * m))
        end = max(start + 6, int(0.60 * m))
        end = min(end, m)
        if end - start < 2:
            return None
    else:
        start, end = fit_range
        if end > m or end - start < 2:
            return None
    return int(start), int(end)


def alpha_from_sv(s: np.ndarray, fit_range=None) -> float:
    """spectral.py:243-273 (np.polyfit degree 1 on log-rank / log-sigma)."""
    s = s[np.isfinite(s) & (s > 0)]
    s = np.sort(s)[::-1]
    m = s.size
    win = alpha_window(m, fit_range)
    if win is None:
        return float("nan")
    start, end = win
    ranks = np.arange(1, m + 1, dtype=np.float64)
    try:
        slope, _ = np.polyfit(np.log(ranks[start:end]), np.log(s[start:end]), 1)
        return float(-slope)
    except Exception:
        return float("nan")


def hill_k(n: int, k=None):
    """Default tail count.  spectral.py:351-353."""
    if k is None:
        k = max(5, int(0.10 * n))
        k = min(k, max(5, n - 1))
    return int(k)


def hill_from_sv(s: np.ndarray, k=None) -> float:
    """spectral.py:343-368."""
    lambdas = (s**2).astype(np.float64)
    lambdas = lambdas[np.isfinite(lambdas) & (lambdas > 0)]
    n = lambdas.size
    if n < 8:
        return float("nan")
    k = hill_k(n, k)
    tail = np.sort(lambdas)[::-1][:k]
    xmin = tail[-1]
    if xmin <= 0 or np.any(tail <= 0):
        return float("nan")
    h = np.log(tail / xmin).mean()
    if h <= 0 or not np.isfinite(h):
        return float("nan")
    return float(1.0 + 1.0 / h)


def _as_f64(w):
    """spectral.py:405-407: anything with .cpu() -> numpy -> float64."""
    if hasattr(w, "cpu"):
        w = w.detach().cpu().numpy() if hasattr(w, "detach") else w.cpu().numpy()
    return np.asarray(w, dtype=np.float64)


def spectral_entropy(w) -> float:
    """spectral.py:49-109."""
    w = np.asarray(w)
    if w.ndim != 2:
        return float("nan")
    s = singular_values(w)
    return float("nan") if s is None else entropy_from_sv(s)


def stable_rank(w) -> float:
    """spectral.py:112-173."""
    w = np.asarray(w)
    if w.ndim != 2:
        return float("nan")
    s = singular_values(w)
    return float("nan") if s is None else stable_rank_from_sv(s)


def alpha_exponent(w, fit_range=None) -> float:
    """spectral.py:176-273."""
    w = np.asarray(w)
    if w.ndim != 2:
        return float("nan")
    s = singular_values(w)
    return float("nan") if s is None else alpha_from_sv(s, fit_range)


def power_law_alpha_hill(w, k=None) -> float:
    """spectral.py:276-368."""
    w = np.asarray(w)
    if w.ndim != 2:
        return float("nan")
    s = singular_values(w)
    return float("nan") if s is None else hill_from_sv(s, k)


def get_spectral_metrics(w) -> dict:
    """spectral.py:371-414: four keys, fixed order."""
    w = _as_f64(w)
    nan = float("nan")
    if w.ndim != 2:
        return dict.fromkeys(METRIC_KEYS, nan)
    s = singular_values(w)
    if s is None:
        return dict.fromkeys(METRIC_KEYS, nan)
    return {
        "spectral_entropy": entropy_from_sv(s),
        "stable_rank": stable_rank_from_sv(s),
        "alpha_exponent": alpha_from_sv(s),
        "pl_alpha_hill": hill_from_sv(s),
    }


def integer_outputs(w, fit_range=None, k=None) -> dict:
    """The integers that decide the estimators:
    m (positive finite SVs), OLS window [start,end), Hill n and k.
    -1 where the reference returns NaN before using the value."""
    out = {"m": 0, "start": -1, "end": -1, "k": -1}
    w = _as_f64(w)
    if w.ndim != 2:
        return out
    s = singular_values(w)
    if s is None:
        return out
    m = int(np.count_nonzero(np.isfinite(s) & (s > 0)))
    out["m"] = m
    win = alpha_window(m, fit_range)
    if win is not None:
        out["start"], out["end"] = win
    lam = s**2
    n = int(np.count_nonzero(np.isfinite(lam) & (lam > 0)))
    if n >= 8:
        out["k"] = hill_k(n, k)
    return out


def aggregate_spectral_metrics(metrics_list) -> dict:
    """spectral.py:417-460: finite-only mean and population std, key order of first dict."""
    if not metrics_list:
        return {}
    result = {}
    for key in metrics_list[0]:
        values = [m[key] for m in metrics_list if np.isfinite(m.get(key, np.nan))]
        if values:
            result[f"{key}_mean"] = float(np.mean(values))
            result[f"{key}_std"] = float(np.std(values))
        else:
            result[f"{key}_mean"] = float("nan")
            result[f"{key}_std"] = float("nan")
    return result


# ------------------------------------------------------------------ distribution
@dataclass
class SpectralDistribution:
    """spectral.py:468-492."""

    name: str
    matrix_type: str
    singular_values: np.ndarray
    eigenvalues: np.ndarray
    normalized_sv: np.ndarray
    cumulative_variance: np.ndarray
    metrics: dict


def get_spectral_distribution(w, name="", matrix_type="unknown"):
    """spectral.py:495-570."""
    w = _as_f64(w)
    if w.ndim != 2:
        return None
    s = singular_values(w)
    if s is None:
        return None
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
        metrics=get_spectral_metrics(w),
    )


# -------------------------------------------------------------------- extraction
@dataclass
class WeightInfo:
    """extraction.py:18-29."""

    name: str
    layer_idx: int | None
    matrix_type: str
    weight: np.ndarray
    shape: tuple


def _layer_idx(name: str):
    """extraction.py:284-290."""
    m = re.search(r"(?:blocks|layers?|encoder\.layer)\.(\d+)", name)
    return int(m.group(1)) if m else None


def _np(t):
    return t.detach().cpu().numpy()


def extract_qkv_weights(model, layer_patterns=None):
    """extraction.py:32-112."""
    out = []
    for name, module in model.named_modules():
        if layer_patterns and not any(p in name for p in layer_patterns):
            continue
        if hasattr(module, "qkv") and hasattr(module.qkv, "weight"):
            qkv = _np(module.qkv.weight)
            d = qkv.shape[1]
            li = _layer_idx(name)
            for tag, blk in (("q", qkv[:d]), ("k", qkv[d : 2 * d]), ("v", qkv[2 * d :])):
                out.append(WeightInfo(f"{name}.qkv.{tag}", li, tag, blk, blk.shape))
        elif hasattr(module, "q_proj") and hasattr(module.q_proj, "weight"):
            li = _layer_idx(name)
            for pn, pt in (("q_proj", "q"), ("k_proj", "k"), ("v_proj", "v")):
                if hasattr(module, pn) and hasattr(getattr(module, pn), "weight"):
                    wt = _np(getattr(module, pn).weight)
                    out.append(WeightInfo(f"{name}.{pn}", li, pt, wt, wt.shape))
    return out


def extract_attention_weights(model, layer_patterns=None):
    """extraction.py:115-155."""
    out = []
    for name, module in model.named_modules():
        if layer_patterns and not any(p in name for p in layer_patterns):
            continue
        if (
            hasattr(module, "proj")
            and hasattr(module.proj, "weight")
            and ("attn" in name.lower() or "attention" in name.lower())
        ):
            wt = _np(module.proj.weight)
            out.append(WeightInfo(f"{name}.proj", _layer_idx(name), "attn_proj", wt, wt.shape))
    return out


def extract_mlp_weights(model, layer_patterns=None):
    """extraction.py:158-205."""
    import torch

    out = []
    for name, module in model.named_modules():
        if layer_patterns and not any(p in name for p in layer_patterns):
            continue
        if (
            ("mlp" in name.lower() or "ffn" in name.lower())
            and hasattr(module, "weight")
            and isinstance(module.weight, torch.Tensor)
        ):
            wt = _np(module.weight)
            last = name.split(".")[-1]
            if "fc1" in name or "0" in last:
                mt = "mlp_up"
            elif "fc2" in name or "2" in last:
                mt = "mlp_down"
            else:
                mt = "mlp"
            out.append(WeightInfo(name, _layer_idx(name), mt, wt, wt.shape))
    return out


def extract_patch_embed_weights(model):
    """extraction.py:208-242."""
    out = []
    for name, module in model.named_modules():
        if "patch_embed" in name.lower() and hasattr(module, "proj") and hasattr(module.proj, "weight"):
            wt = _np(module.proj.weight)
            if wt.ndim == 4:
                wt = wt.reshape(wt.shape[0], -1)
            out.append(WeightInfo(f"{name}.proj", None, "patch_embed", wt, wt.shape))
    return out


def extract_all_weights(
    model, layer_patterns=None, include_qkv=True, include_proj=True, include_mlp=False, include_patch_embed=True
):
    """extraction.py:245-281."""
    out = []
    if include_qkv:
        out += extract_qkv_weights(model, layer_patterns)
    if include_proj:
        out += extract_attention_weights(model, layer_patterns)
    if include_mlp:
        out += extract_mlp_weights(model, layer_patterns)
    if include_patch_embed:
        out += extract_patch_embed_weights(model)
    return out


# ------------------------------------------------------------------------ callers
def extract_and_analyze_weights(model, device=None) -> dict:
    """run_spectral_analysis.py:297-345."""
    model.eval()
    all_w = extract_qkv_weights(model) + extract_attention_weights(model) + extract_mlp_weights(model)
    per_layer, svs, mlist = {}, {}, []
    for wi in all_w:
        m = get_spectral_metrics(wi.weight)
        per_layer[wi.name] = m
        mlist.append(m)
        s = singular_values(wi.weight.astype(np.float64))
        svs[wi.name] = [] if s is None else s.tolist()
    return {
        "per_layer_metrics": per_layer,
        "aggregated_metrics": aggregate_spectral_metrics(mlist),
        "singular_values": svs,
    }


def compute_spectral_metrics_trainer(
    model, layer_patterns=None, extract_qkv=True, extract_mlp=False, extract_patch_embed=True
) -> dict:
    """training/base.py:379-416 (overall + per matrix_type aggregates)."""
    model.eval()
    weights = extract_all_weights(
        model,
        layer_patterns=layer_patterns,
        include_qkv=extract_qkv,
        include_mlp=extract_mlp,
        include_patch_embed=extract_patch_embed,
    )
    if not weights:
        return {}
    all_m, by_type = [], {}
    for w in weights:
        m = get_spectral_metrics(w.weight)
        all_m.append(m)
        by_type.setdefault(w.matrix_type, []).append(m)
    result = aggregate_spectral_metrics(all_m)
    for mt, ms in by_type.items():
        for k, v in aggregate_spectral_metrics(ms).items():
            result[f"{mt}_{k}"] = v
    return result


# ------------------------------------------------------- cpu_baseline cost model
def reference_cost_metrics(w) -> dict:
    """What the reference actually *executes* per matrix in the six-scenario
    driver: four independent SVDs inside get_spectral_metrics (spectral.py:409-414)
    plus a fifth for the SV list (run_spectral_analysis.py:331-333).  Used only by
    bench.py's cpu_baseline / --impl reference legs so the CPU arm is charged the
    reference's real work, not the oracle's shared-SVD shortcut."""
    w = _as_f64(w)
    out = {
        "spectral_entropy": spectral_entropy(w),
        "stable_rank": stable_rank(w),
        "alpha_exponent": alpha_exponent(w),
        "pl_alpha_hill": power_law_alpha_hill(w),
    }
    s = singular_values(w)
    return {"metrics": out, "sv": s}


reference_cost_svds = 5


# -------------------------------------------------------------------- Clauset x_min scan (extra output)
def clauset_xmin_scan(w):
    """Clauset-Shalizi-Newman (SIAM Review 51, 2009, section 3.3) x_min scan on the eigenvalue spectrum lambda = sigma^2
    of `w`: for every candidate cutoff x_min = lambda_(k) (at least two points in the tail) the continuous MLE
    alpha = 1 + t / sum ln(lambda_i / x_min) (their eq. 3.1) and the Kolmogorov-Smirnov distance (eq. 3.9) between the
    tail's empirical CDF and the fitted P(x) = 1 - (x / x_min)^(1 - alpha); the cutoff with the smallest distance wins
    (ties: the smaller cutoff).  The REFERENCE HAS NO SUCH SCAN (its alpha estimators are the fixed-window
    OLS slope and the fixed-k Hill estimator): this restates the published algorithm and is the oracle of the opt-in
    `clauset` output only -- parity unpinned against the reference by construction."""
    s = singular_values(w)
    nan = float("nan")
    out = {"alpha": nan, "xmin": nan, "ks_distance": nan, "xmin_index": -1, "tail_count": -1}
    if s is None:
        return out
    lam = np.sort(np.asarray(s, dtype=np.float64) ** 2)
    lam = lam[np.isfinite(lam) & (lam > 0)]
    m = lam.size
    if m < 8:
        return out
    ll = np.log(lam)
    best = None
    for k in range(m - 1):
        t = m - k
        sl = float(np.sum(ll[k:] - ll[k]))
        if not sl > 0.0:
            continue
        a = 1.0 + t / sl
        P = 1.0 - np.exp((1.0 - a) * (ll[k:] - ll[k]))
        j = np.arange(t, dtype=np.float64)
        D = float(np.max(np.maximum(np.abs((j + 1) / t - P), np.abs(j / t - P))))
        if best is None or D < best[0]:
            best = (D, a, k)
    if best is None:
        return out
    D, a, k = best
    return {"alpha": a, "xmin": float(lam[k]), "ks_distance": D, "xmin_index": m - 1 - k, "tail_count": m - k}


# -------------------------------------------------------------------- singular-vector consumers (8f rank 4)
def truncate_weight_matrix(weight, retention_ratio=0.9, min_rank=1):
    """metrics/tail_truncation.py:63-105, restated (full SVD with vectors, SciPy)."""
    from scipy.linalg import svd

    U, s, Vt = svd(np.asarray(weight).astype(np.float64), full_matrices=False)
    k = min(max(min_rank, int(np.ceil(len(s) * retention_ratio))), len(s))
    st = s.copy()
    st[k:] = 0.0
    total = np.sum(s**2)
    return (U @ np.diag(st) @ Vt).astype(np.asarray(weight).dtype), {
        "original_rank": int(np.sum(s > 1e-10)), "truncated_rank": k, "energy_retained": float(np.sum(st**2) / total) if total > 0 else 1.0}


def truncate_by_energy(weight, energy_threshold=0.99, min_rank=1):
    """metrics/tail_truncation.py:108-152, restated."""
    from scipy.linalg import svd

    U, s, Vt = svd(np.asarray(weight).astype(np.float64), full_matrices=False)
    total = np.sum(s**2)
    if total <= 0:
        return weight, {"original_rank": 0, "truncated_rank": 0, "energy_retained": 1.0}
    k = int(np.searchsorted(np.cumsum(s**2) / total, energy_threshold) + 1)
    k = max(min_rank, min(k, len(s)))
    st = s.copy()
    st[k:] = 0.0
    return (U @ np.diag(st) @ Vt).astype(np.asarray(weight).dtype), {
        "original_rank": int(np.sum(s > 1e-10)), "truncated_rank": int(k), "energy_retained": float(np.sum(st**2) / total)}


def compute_rank_reducing_gradient(weight):
    """metrics/gradient_alignment.py:48-70, restated: U @ Vt."""
    from scipy.linalg import svd

    U, s, Vt = svd(np.asarray(weight).astype(np.float64), full_matrices=False)
    return U @ Vt
