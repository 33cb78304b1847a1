"""Singular-vector consumers on the device: tail truncation and the rank-reducing gradient U V^T as
Newton-Schulz matrix functions on the FP64 tensor cores, against the oracle restatements of the reference's SVD-based
functions (which tests/test_oracle.py pins to outputs of the real reference)."""

import numpy as np
import pytest
import spectral_oracle as orc
import torch
from _inputs import build_case, trunc_normal

pytestmark = pytest.mark.gpu


def test_dgemm_dmma_against_float64_matmul():
    from vision_spectra_b200.lowrank import dgemm

    g = torch.Generator(device="cuda").manual_seed(0)
    for (m, k, n, ta, tb) in [(64, 64, 64, False, False), (100, 37, 75, False, False), (192, 768, 192, True, False),
                              (33, 200, 9, False, True), (130, 70, 130, True, True), (1, 5, 1, False, False)]:
        a = torch.randn((k, m) if ta else (m, k), generator=g, device="cuda", dtype=torch.float64)
        b = torch.randn((n, k) if tb else (k, n), generator=g, device="cuda", dtype=torch.float64)
        c = torch.randn((m, n), generator=g, device="cuda", dtype=torch.float64)
        ref = 0.7 * (a.T if ta else a) @ (b.T if tb else b) - 1.3 * c + 0.25 * torch.eye(m, n, device="cuda", dtype=torch.float64)
        dgemm(a, b, c, alpha=0.7, beta=-1.3, gamma=0.25, trans_a=ta, trans_b=tb)
        torch.cuda.synchronize()
        assert torch.max(torch.abs(c - ref)).item() < 1e-12 * max(1.0, k), (m, k, n, ta, tb)


CASES = ["vit:C:0:q", "vit:C:0:mlp_up", "vit:E:0:mlp_down", "randn:30x50:f64", "sgd:96x384", "powerlaw:64:0.5:f32", "vit:A:0:q", "vit:A:0:mlp_down"]


@pytest.mark.parametrize("name", CASES)
def test_tail_truncation_matches_the_svd_route(name):
    from vision_spectra_b200.metrics.tail_truncation import truncate_by_energy, truncate_weight_matrix

    w = build_case(name)
    nrm = np.linalg.norm(w.astype(np.float64))
    for fn, ofn, arg in ((truncate_weight_matrix, orc.truncate_weight_matrix, 0.9), (truncate_weight_matrix, orc.truncate_weight_matrix, 0.5),
                         (truncate_by_energy, orc.truncate_by_energy, 0.95)):
        got, info = fn(w, arg)
        ref, rinfo = ofn(w, arg)
        assert got.dtype == w.dtype and got.shape == w.shape
        assert (info["original_rank"], info["truncated_rank"]) == (rinfo["original_rank"], rinfo["truncated_rank"])
        assert abs(info["energy_retained"] - rinfo["energy_retained"]) < 1e-9
        tol = 2e-6 if w.dtype == np.float32 else 1e-9  # the float32 cast of the result dominates for fp32 inputs
        assert np.linalg.norm(got.astype(np.float64) - ref.astype(np.float64)) / nrm < tol, (name, arg)
    # torch tensor in -> tensor out, on the device
    t, _ = truncate_weight_matrix(torch.from_numpy(w).cuda(), 0.9)
    assert t.is_cuda and t.dtype == torch.from_numpy(w).dtype


@pytest.mark.parametrize("name", CASES)
def test_rank_reducing_gradient_and_alignment(name):
    from vision_spectra_b200.metrics.gradient_alignment import compute_gradient_alignment, compute_rank_reducing_gradient

    w = build_case(name)
    got = compute_rank_reducing_gradient(w)
    ref = orc.compute_rank_reducing_gradient(w)
    assert got.shape == w.shape and got.dtype == np.float64
    assert np.max(np.abs(got - ref)) < 1e-7, (name, np.max(np.abs(got - ref)))
    # partial isometry: all singular values of U V^T are one
    s = np.linalg.svd(got, compute_uv=False)
    assert np.max(np.abs(s - 1.0)) < 1e-9
    rng = np.random.default_rng(3)
    grad = trunc_normal(rng, w.shape)
    res = compute_gradient_alignment(grad, w)
    rf, gf = ref.flatten(), grad.flatten().astype(np.float64)
    cos = float(np.dot(gf, rf) / (np.linalg.norm(gf) * np.linalg.norm(rf)))
    assert abs(res.cosine_similarity - cos) < 1e-7 and res.is_aligned == (cos > 0)
    assert abs(res.rank_reducing_grad_norm - np.linalg.norm(rf)) < 1e-6
