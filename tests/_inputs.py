"""Seeded input builders shared by oracle/gen_golden.py and the tests.

Every case is rebuilt from its name alone (NumPy PCG64 streams are stable), so
the golden files only have to hold the reference's OUTPUTS plus an input
checksum that guards against generator drift.
"""

from __future__ import annotations

import zlib

import numpy as np

# d, depth for the BASELINE.json configs
VIT_CONFIGS = {"E": (32, 1), "C": (96, 3), "A": (192, 6), "Base": (768, 12)}


def _seed(name: str) -> int:
    return zlib.crc32(name.encode()) & 0x7FFFFFFF


def trunc_normal(rng, shape, std=0.02, a=-2.0, b=2.0) -> np.ndarray:
    """timm's default Linear init: N(0, std^2) truncated to [a, b] (absolute
    bounds, i.e. effectively untruncated at std=0.02).  SURVEY 8d."""
    x = rng.standard_normal(shape) * std
    bad = (x < a) | (x > b)
    while bad.any():
        x[bad] = rng.standard_normal(int(bad.sum())) * std
        bad = (x < a) | (x > b)
    return x.astype(np.float32)


def vit_block_matrices(d: int, rng) -> list[tuple[str, np.ndarray]]:
    """The six matrices the driver analyses per transformer block
    (run_spectral_analysis.py:313-317): q,k,v as row-blocks of one fused
    [3d,d] buffer, proj [d,d], fc1 [4d,d], fc2 [d,4d]."""
    qkv = trunc_normal(rng, (3 * d, d))
    proj = trunc_normal(rng, (d, d))
    fc1 = trunc_normal(rng, (4 * d, d))
    fc2 = trunc_normal(rng, (d, 4 * d))
    return [
        ("q", qkv[:d]),
        ("k", qkv[d : 2 * d]),
        ("v", qkv[2 * d :]),
        ("attn_proj", proj),
        ("mlp_up", fc1),
        ("mlp_down", fc2),
    ]


def power_law_matrix(n: int, alpha: float, rng, dtype) -> np.ndarray:
    """U diag(i^-alpha) V^T as in tests/test_metrics.py:113-135."""
    u = np.linalg.qr(rng.standard_normal((n, n)))[0]
    v = np.linalg.qr(rng.standard_normal((n, n)))[0]
    s = np.arange(1, n + 1, dtype=np.float64) ** (-alpha)
    return (u @ np.diag(s) @ v).astype(dtype)


def build_case(name: str) -> np.ndarray:
    """name -> input array.  Names are the keys of the golden file."""
    rng = np.random.default_rng(_seed(name))
    kind, _, rest = name.partition(":")
    if kind == "vit":  # vit:<cfg>:<block>:<which>
        cfg, blk, which = rest.split(":")
        d, _ = VIT_CONFIGS[cfg]
        rng = np.random.default_rng(_seed(f"vit:{cfg}:{blk}"))
        return dict(vit_block_matrices(d, rng))[which]
    if kind == "powerlaw":  # powerlaw:<n>:<alpha>:<f32|f64>
        n, alpha, dt = rest.split(":")
        return power_law_matrix(int(n), float(alpha), rng, np.float32 if dt == "f32" else np.float64)
    if kind == "randn":  # randn:<r>x<c>:<f32|f64>[:scale]
        parts = rest.split(":")
        r, c = (int(v) for v in parts[0].split("x"))
        dt = np.float32 if parts[1] == "f32" else np.float64
        scale = float(parts[2]) if len(parts) > 2 else 1.0
        return (rng.standard_normal((r, c)) * scale).astype(dt)
    if kind == "eye":
        return np.eye(int(rest), dtype=np.float32)
    if kind == "zeros":
        r, c = (int(v) for v in rest.split("x"))
        return np.zeros((r, c), dtype=np.float32)
    if kind == "rank1":
        n = int(rest)
        u = rng.standard_normal((n, 1))
        return (u @ u.T).astype(np.float32)
    if kind == "nan":
        w = rng.standard_normal((16, 16)).astype(np.float32)
        w[3, 5] = np.nan
        return w
    if kind == "inf":
        w = rng.standard_normal((16, 16)).astype(np.float32)
        w[0, 0] = np.inf
        return w
    if kind == "illcond":  # tests/test_metrics.py:279-293 (kappa ~ 1e10), f64
        n = int(rest)
        u = rng.standard_normal((n, n))
        v = rng.standard_normal((n, n))
        return u @ np.diag(np.logspace(0, -10, n)) @ v
    if kind == "sgd":  # a random-init matrix pushed toward low rank (trained-like spectrum)
        r, c = (int(v) for v in rest.split("x"))
        base = trunc_normal(rng, (r, c)).astype(np.float64)
        k = max(2, min(r, c) // 16)
        a = rng.standard_normal((r, k)) * 0.05
        b = rng.standard_normal((k, c)) * 0.05
        return (base + a @ b).astype(np.float32)
    if kind == "vec":
        return rng.standard_normal(int(rest)).astype(np.float32)
    raise KeyError(name)


def checksum(x: np.ndarray) -> str:
    return f"{zlib.crc32(np.ascontiguousarray(x).tobytes()):08x}"


def golden_case_names() -> list[str]:
    names = []
    for cfg in ("E", "C", "A"):
        for which in ("q", "k", "v", "attn_proj", "mlp_up", "mlp_down"):
            names.append(f"vit:{cfg}:0:{which}")
    names += ["vit:A:1:q", "vit:A:1:mlp_down", "vit:C:2:attn_proj"]
    names += ["vit:Base:0:q", "vit:Base:0:mlp_up", "vit:Base:0:mlp_down"]
    for alpha in ("0.5", "1.0", "2.0", "4.0"):
        names.append(f"powerlaw:100:{alpha}:f64")
    names += ["powerlaw:100:1.0:f32", "powerlaw:64:0.5:f32"]
    names += [
        "randn:64x64:f64",
        "randn:50x50:f64",
        "randn:100x100:f64",
        "randn:30x50:f64",
        "randn:32x128:f32",
        "randn:7x7:f32",
        "randn:8x8:f32",
        "randn:4x100:f32",
        "randn:100x4:f32",
        "randn:9x33:f32",
        "randn:1x1:f32",
        "randn:32x32:f32:1e-10",
        "randn:32x32:f32:1e6",
        "randn:32x32:f64:1e-30",
        "randn:200x130:f32",
        "randn:257x65:f32",
        "eye:10",
        "eye:50",
        "zeros:12x12",
        "rank1:10",
        "nan:0",
        "inf:0",
        "illcond:50",
        "sgd:192x192",
        "sgd:768x192",
        "sgd:96x384",
        "vec:10",
    ]
    return names
