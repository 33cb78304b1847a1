"""The C-ABI library loads and exports every symbol include/vspectra.h declares
(no compute calls: this runs without a GPU)."""

import ctypes
import re
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]


@pytest.fixture(scope="module")
def lib():
    from vision_spectra_b200 import _native as nat
    from vision_spectra_b200.build import build_native

    build_native()  # no-op when lib/libvspectra.so is up to date
    return nat.load()


def _declared_functions():
    text = (ROOT / "include" / "vspectra.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vsp_[a-z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported(lib):
    from vision_spectra_b200 import _native as nat

    declared = _declared_functions()
    assert len(declared) >= 16
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/vspectra.h but not exported"
    assert sorted(nat.EXPORTED_SYMBOLS) == declared


def test_version_and_error_strings(lib):
    assert lib.vsp_version() == 100
    assert lib.vsp_error_string(0) == b"ok"
    assert b"argument" in lib.vsp_error_string(-1)
    assert b"workspace" in lib.vsp_error_string(-3)


def test_record_layout_matches_header():
    from vision_spectra_b200 import _native as nat

    assert nat.RECORD_DTYPE.itemsize == 64
    assert nat.RECORD_DTYPE.fields["metrics"][1] == 32
    assert ctypes.sizeof(nat.VspOpts) == 32


def test_shape_validation_and_layout_queries(lib):
    """Host-only entry points: argument validation, workspace size, SV offsets."""
    from vision_spectra_b200 import _native as nat

    rows, cols = nat.i32([32, 128, 32, 768]), nat.i32([32, 32, 128, 3072])
    ws = lib.vsp_workspace_bytes(4, nat.p32(rows), nat.p32(cols))
    # the bound is laid out by the plan code itself (GPU test test_workspace_bound_covers_every_plan holds it against
    # real plans); here: it covers the FP64 Gram triangles, the six digit planes and one FP64 copy of the largest
    # matrix per shape class (re-solve pool), and stays within a few pages of alignment of their sum
    tri = sum(n * (n + 1) // 2 + 2 * n for n in (32, 32, 32, 768)) * 8
    planes = sum(6 * n * k for n, k in ((32, 32), (32, 128), (32, 128), (768, 3072)))
    pool = 3 * 128 * 32 * 8 + 768 * 3072 * 8
    # (+ band outputs, exchange buffers of the cluster re-solve, rounding flags, alignment: within 10 % + 1 MB)
    assert tri + planes + pool <= ws <= 1.1 * (tri + planes + pool) + (1 << 20)
    offs = np.zeros(5, np.int64)
    assert lib.vsp_sv_offsets(4, nat.p32(rows), nat.p32(cols), nat.p64(offs)) == 0
    assert offs.tolist() == [0, 32, 64, 96, 864]
    assert lib.vsp_workspace_bytes(-1, nat.p32(rows), nat.p32(cols)) == -1
    assert lib.vsp_workspace_bytes(4, nat.p32(nat.i32([0, 1, 1, 1])), nat.p32(cols)) == -1
    assert lib.vsp_workspace_bytes(1, nat.p32(nat.i32([5000])), nat.p32(nat.i32([6000]))) == -2
    assert 0 < lib.vsp_workspace_bytes(0, None, None) <= 8192


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import vision_spectra_b200 as pkg
    from vision_spectra_b200 import _native as nat
    from vision_spectra_b200.metrics import get_spectral_metrics

    with pytest.raises(nat.NativeError):
        pkg.SpectraEngine()
    with pytest.raises(nat.NativeError):
        get_spectral_metrics(np.eye(8, dtype=np.float32))
    # data-level failure is still NaN, not an exception (reference spectral.py:87)
    assert all(np.isnan(v) for v in get_spectral_metrics(np.zeros(5)).values())


def test_product_never_imports_oracle():
    """Guard for the rule that only tests/, smoke() and bench.py may touch oracle/."""
    pkg_dir = ROOT / "vision-spectra_b200"
    bad = re.compile(r"^\s*(?:from|import)\s+(?:scipy|spectral_oracle|oracle)\b|oracle[/\\]|#include\s+\"[^\"]*oracle", re.M)
    for path in list(pkg_dir.rglob("*.py")) + list(pkg_dir.rglob("*.cu")) + list(pkg_dir.rglob("*.cuh")):
        assert not bad.search(path.read_text()), path


def test_torch_extension_loads_and_registers_the_op():
    """lib/vspectra_torch.so (csrc/torch_ext.cpp) loads without a GPU and registers vision_spectra_b200::analyze_batch
    with the documented schema; calling it with CPU tensors is an error (no fallback kernel is registered)."""
    import pytest
    import torch
    from vision_spectra_b200 import _native as nat

    ops = nat.load_torch_ext()
    schema = str(torch.ops.vision_spectra_b200.analyze_batch.default._schema)
    assert "Tensor[] matrices" in schema and "int hill_k=-1" in schema and "bool want_sv=True" in schema and "int dist_k=0" in schema and "bool clauset=False" in schema
    with pytest.raises((RuntimeError, NotImplementedError)):
        ops.analyze_batch([torch.zeros(4, 4)], -1, -1, -1, True, 0, False)
