"""Singular-vector consumers without singular vectors.

The reference's tail truncation (`metrics/tail_truncation.py:63-152`: U diag(s_k) V^T with the tail singular values
zeroed) and rank-reducing gradient (`metrics/gradient_alignment.py:48-70`: U V^T, the gradient of the nuclear norm) both
take a full LAPACK SVD with vectors on the CPU.  Both are matrix FUNCTIONS of W:

    truncation   W_k = W P  (or P W),  P = (I + sign(G - mu I)) / 2   the spectral projector of the smaller Gram matrix
                                                                     G onto the eigenvalues above the cut mu
    polar factor U V^T = W (W^T W)^(-1/2)

and both functions come out of Newton-Schulz iterations (X <- X (3 I - X^2) / 2 for the sign function, X <- X (3 I -
X^T X) / 2 for the polar factor) that need nothing but matrix products -- the B200-native formulation SURVEY names.
The products run on the FP64 tensor cores (csrc/dgemm_dmma.cuh behind `vsp_dgemm_batched`); the spectrum that fixes the
cut, the scaling and the NUMBER of iterations (no convergence test, no host synchronisation inside the loop) comes from
the hot path (`SpectraEngine.analyze`).  PyTorch: memory, dtype casts and streams only.
"""

from __future__ import annotations

import math
from typing import Any

import numpy as np
import torch

from . import _native as nat
from .engine import SpectraEngine, default_engine

_MAX_ITERS = 80


def _ptrs(t: torch.Tensor) -> torch.Tensor:
    return torch.tensor([t.data_ptr()], dtype=torch.int64, device=t.device)


def dgemm(a: torch.Tensor, b: torch.Tensor, c: torch.Tensor, alpha=1.0, beta=0.0, gamma=0.0, trans_a=False, trans_b=False) -> torch.Tensor:
    """c = alpha op(a) op(b) + beta c + gamma I on the device (float64, row-major 2-D tensors with unit column stride)."""
    lib = nat.load()
    for x in (a, b, c):
        if x.dtype != torch.float64 or x.ndim != 2 or not x.is_cuda or x.stride(1) != 1:
            raise ValueError("dgemm: float64 2-D CUDA tensors with unit column stride")
    m, k = (a.shape[1], a.shape[0]) if trans_a else a.shape
    k2, n = (b.shape[1], b.shape[0]) if trans_b else b.shape
    if k != k2 or tuple(c.shape) != (m, n):
        raise ValueError("dgemm: shapes do not match")
    stream = torch.cuda.current_stream(a.device).cuda_stream
    pa, pb, pc = _ptrs(a), _ptrs(b), _ptrs(c)
    with torch.cuda.device(a.device):
        nat.check(lib.vsp_dgemm_batched(1, m, n, k, float(alpha), pa.data_ptr(), a.stride(0), int(trans_a), pb.data_ptr(), b.stride(0),
                                        int(trans_b), float(beta), float(gamma), pc.data_ptr(), c.stride(0), stream), "vsp_dgemm_batched")
    return c


def _iterations(smallest: float) -> int:
    """Newton-Schulz steps until a scaled value `smallest` in (0, 1] has reached 1 to working precision: it grows by
    3/2 per step while small, then converges quadratically (six more steps cover 0.5 -> 1 - 1e-16)."""
    if not (smallest > 0.0):
        return _MAX_ITERS
    return min(_MAX_ITERS, int(math.ceil(math.log(1.0 / min(smallest, 1.0)) / math.log(1.5))) + 7)


def _device_f64(w: Any, engine: SpectraEngine) -> torch.Tensor:
    t = w.detach() if isinstance(w, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(w))
    return t.to(device=engine.device, dtype=torch.float64).contiguous()


def polar_factor(weight: Any, engine: SpectraEngine | None = None) -> torch.Tensor:
    """U V^T of `weight` (float64, on the device): X_0 = W / sigma_max, X <- X (3 I - X^T X) / 2 on the smaller side."""
    engine = engine or default_engine(weight.device if isinstance(weight, torch.Tensor) and weight.is_cuda else None)
    w = _device_f64(weight, engine)
    _, svs, _ = engine.analyze([w])
    s = svs[0]
    if s is None or not (s[0] > 0):
        return torch.zeros_like(w)
    pos = s[s > s[0] * 1e-15]
    x = w / float(s[0])
    tall = w.shape[0] >= w.shape[1]
    n = min(w.shape)
    z = torch.empty((n, n), dtype=torch.float64, device=w.device)
    y = torch.empty_like(x)
    for _ in range(_iterations(float(pos[-1] / s[0]))):
        if tall:
            dgemm(x, x, z, alpha=-0.5, gamma=1.5, trans_a=True)  # 1.5 I - 0.5 X^T X
            dgemm(x, z, y)
        else:
            dgemm(x, x, z, alpha=-0.5, gamma=1.5, trans_b=True)  # 1.5 I - 0.5 X X^T
            dgemm(z, x, y)
        x, y = y, x
    return x


def spectral_truncation(weight: Any, keep: int, engine: SpectraEngine | None = None, singular_values: np.ndarray | None = None) -> torch.Tensor:
    """W with all but its `keep` largest singular values zeroed (float64, on the device): W P or P W with the spectral
    projector of the smaller Gram matrix, P = (I + sign(G - mu I)) / 2, mu halfway between the squares of the last
    kept and the first dropped singular value."""
    engine = engine or default_engine(weight.device if isinstance(weight, torch.Tensor) and weight.is_cuda else None)
    w = _device_f64(weight, engine)
    s = singular_values
    if s is None:
        _, svs, _ = engine.analyze([w])
        s = svs[0]
    n = min(w.shape)
    if s is None or keep >= n:
        return w.clone()
    if keep <= 0:
        return torch.zeros_like(w)
    lam = np.asarray(s, dtype=np.float64) ** 2
    mu = 0.5 * (lam[keep - 1] + lam[keep])
    rho = max(lam[0] - mu, mu - lam[-1])  # spectral radius of G - mu I
    gap = min(lam[keep - 1] - mu, mu - lam[keep])
    if not (gap > 0.0) or not (rho > 0.0):
        return w.clone()  # the cut falls inside a multiple singular value: nothing well defined to drop
    tall = w.shape[0] >= w.shape[1]
    g = torch.empty((n, n), dtype=torch.float64, device=w.device)
    # S_0 = (G - mu I) / rho
    dgemm(w, w, g, alpha=1.0 / rho, gamma=-mu / rho, trans_a=tall, trans_b=not tall)
    z = torch.empty_like(g)
    y = torch.empty_like(g)
    for _ in range(_iterations(gap / rho)):
        dgemm(g, g, z, alpha=-0.5, gamma=1.5)  # 1.5 I - 0.5 S^2
        dgemm(g, z, y)
        g, y = y, g
    # P = (I + sign) / 2, folded into the last product: W P = 0.5 W S + 0.5 W
    out = w.clone()
    if tall:
        dgemm(w, g, out, alpha=0.5, beta=0.5)
    else:
        dgemm(g, w, out, alpha=0.5, beta=0.5)
    return out
