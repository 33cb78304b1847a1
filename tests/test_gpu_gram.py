"""Stage 1 in isolation: the tcgen05 int8-split Gram (fp32 inputs) and the FP64 CUDA-core
Gram (fp64 inputs) against torch's float64 matmul.  The split is exact up to 2^-46, so the
tolerance is FP64-level (1e-13 relative to sqrt(G_ii G_jj))."""

import ctypes

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _gram_via_abi(mats):
    from vision_spectra_b200 import _native as nat

    lib = nat.load()
    dev = mats[0].device
    count = len(mats)
    rows, cols = nat.i32([m.shape[0] for m in mats]), nat.i32([m.shape[1] for m in mats])
    ld = nat.i64([m.stride(0) for m in mats])
    dtype = nat.VSP_F32 if mats[0].dtype == torch.float32 else nat.VSP_F64
    plan = ctypes.c_void_p()
    nat.check(lib.vsp_plan_create(count, nat.p32(rows), nat.p32(cols), nat.p64(ld), dtype, None, ctypes.byref(plan)))
    ws = torch.empty(int(lib.vsp_plan_workspace_bytes(plan)), dtype=torch.uint8, device=dev)
    ns = np.minimum(rows, cols).astype(np.int64)
    out = torch.full((int((ns * ns).sum()),), float("nan"), dtype=torch.float64, device=dev)
    ptrs = (ctypes.c_void_p * count)(*[m.data_ptr() for m in mats])
    nat.check(lib.vsp_plan_debug_gram(plan, ptrs, out.data_ptr(), ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    lib.vsp_plan_destroy(plan)
    res, off = [], 0
    for n in ns:
        res.append(out[off : off + n * n].view(n, n).cpu().numpy())
        off += n * n
    return res


def _ref_gram(m):
    x = m.double()
    return (x @ x.T if m.shape[0] <= m.shape[1] else x.T @ x).cpu().numpy()


@pytest.mark.parametrize("dtype", [torch.float32, torch.float64])
def test_gram_matches_float64_matmul(dtype):
    g = torch.Generator(device="cuda").manual_seed(1)
    shapes = [(32, 32), (128, 32), (32, 128), (96, 96), (384, 96), (96, 384), (192, 192), (768, 192), (192, 768),
              (7, 7), (9, 33), (257, 65), (200, 130), (1, 1), (100, 4), (300, 300),
              # last 128-row tile of at most 64 rows (transposed-tile enumeration, gram_i8.cuh: I8Class::xt) and just above
              (320, 320), (500, 448), (450, 450), (193, 700)]
    mats = [(torch.randn(s, generator=g, device="cuda", dtype=torch.float32) * 0.02).to(dtype) for s in shapes]
    # rows spanning 9 decades, denormals, an exact-zero row
    wide = torch.randn(64, 256, generator=g, device="cuda") * torch.logspace(-6, 3, 64, device="cuda")[:, None]
    wide[5] = 0
    tiny = torch.randn(40, 40, generator=g, device="cuda") * 1e-41
    mats += [wide.to(dtype), tiny.float().to(dtype)]
    got = _gram_via_abi(mats)
    for m, G in zip(mats, got):
        ref = _ref_gram(m)
        dg = np.sqrt(np.maximum(np.diag(ref), 0))
        scale = np.outer(dg, dg)
        scale[scale == 0] = 1.0
        err = np.max(np.abs(G - ref) / scale)
        assert np.all(np.isfinite(G)) and err < 1e-13, (tuple(m.shape), err)
        assert np.array_equal(G, G.T)


def test_gram_random_orders():
    """Seeded sweep over Gram orders around the tile boundaries (64 / 128 / 192 / 256, multiples of 8 and not), both
    orientations, short and long contractions: every tile enumeration, frame offset and store path of the int8 kernel."""
    rng = np.random.default_rng(7)
    g = torch.Generator(device="cuda").manual_seed(7)
    shapes = []
    for n in list(rng.integers(57, 72, 4)) + list(rng.integers(120, 140, 6)) + list(rng.integers(185, 201, 6)) + list(
            rng.integers(250, 330, 6)) + list(rng.integers(380, 460, 3)):
        K = int(rng.choice([int(n), int(n) + 5, 2 * int(n), 777]))
        K = max(K, int(n))
        shapes.append((int(n), K) if rng.random() < 0.5 else (K, int(n)))
    mats = [torch.randn(s, generator=g, device="cuda", dtype=torch.float32) * 0.02 for s in shapes]
    got = _gram_via_abi(mats)
    for m, G in zip(mats, got):
        ref = _ref_gram(m)
        dg = np.sqrt(np.maximum(np.diag(ref), 0))
        err = np.max(np.abs(G - ref) / np.outer(dg, dg))
        assert np.all(np.isfinite(G)) and err < 1e-13, (tuple(m.shape), err)
        assert np.array_equal(G, G.T)


def test_gram_views_and_nonfinite():
    g = torch.Generator(device="cuda").manual_seed(2)
    qkv = torch.randn(3 * 192, 192, generator=g, device="cuda") * 0.02
    views = [qkv[:192], qkv[192:384], qkv[384:]]
    col = (torch.randn(96, 200, generator=g, device="cuda") * 0.02)[:, 10:106]  # ld = 200
    bad = torch.randn(48, 48, generator=g, device="cuda")
    bad[3, 7] = float("inf")
    got = _gram_via_abi(views + [col, bad])
    for m, G in zip(views + [col], got[:4]):
        ref = _ref_gram(m)
        assert np.max(np.abs(G - ref)) / np.max(np.abs(ref)) < 1e-13
    assert not np.isfinite(got[4][3, 3])  # the poisoned Gram row is flagged on the diagonal


def test_gram_within_row_dynamic_range():
    """The int8 split anchors every Gram row at its largest exponent and keeps 42 bits below it: elements more than 2^17
    below the row maximum are ROUNDED at 2^-41 of it (gram_i8.cuh).  Outlier columns and single huge entries are the
    cases where that happens; the Gram matrix must then still be accurate to ~K 2^-42 relative to sqrt(G_ii G_jj)
    (the eigensolve compensates by re-solving such matrices from W earlier: kRefineRatioInexact), and stay exact
    (1e-13) when the outliers are within 2^17."""
    g = torch.Generator(device="cuda").manual_seed(3)
    base = lambda r, c: torch.randn(r, c, generator=g, device="cuda") * 0.02
    far_cols = base(192, 192)
    far_cols[:, [3, 77, 150]] *= 1e6  # three outlier columns, 2^20 above the rest
    far_entry = base(96, 384)
    far_entry[10, 20] = 3.0e4  # one entry 2^20 above its row
    far_entry[50, 7] = -1.0e5
    tall = base(768, 192)
    tall[:, 5] *= 1e6  # tall: the Gram index is the column -> a whole Gram row is large: exact again
    near_cols = base(192, 192)
    near_cols[:, [1, 100]] *= 5e4  # 2^15.6: inside the exact range
    got = _gram_via_abi([far_cols, far_entry, tall, near_cols])
    for m, G, tol in zip([far_cols, far_entry, tall, near_cols], got, [1e-10, 1e-10, 1e-13, 1e-12]):
        ref = _ref_gram(m)
        dg = np.sqrt(np.maximum(np.diag(ref), 0))
        err = np.max(np.abs(G - ref) / np.outer(dg, dg))
        assert np.all(np.isfinite(G)) and err < tol, (tuple(m.shape), err, tol)
