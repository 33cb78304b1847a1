"""ctypes binding of the C-ABI in include/vspectra.h.

The CUDA library is the only implementation: if it is missing this module raises,
it never falls back to a CPU path.
"""

from __future__ import annotations

import ctypes
from ctypes import POINTER, c_char_p, c_double, c_int32, c_int64, c_void_p
from pathlib import Path

import numpy as np

import os

# VSP_LIB: development override (an alternative build of the same library, e.g. -DVSP_PHASE_TIMING)
LIB_PATH = Path(os.environ.get("VSP_LIB") or Path(__file__).resolve().parent / "lib" / "libvspectra.so")

VSP_F32, VSP_F64 = 0, 1
ST_NONFINITE, ST_ZERO, ST_FEW_SV, ST_ALPHA_NAN, ST_HILL_NAN, ST_REFINED, ST_ILLCOND = 1, 2, 4, 8, 16, 32, 64

# vsp_record, 64 bytes
RECORD_DTYPE = np.dtype(
    [
        ("item", "<i4"),
        ("status", "<i4"),
        ("m", "<i4"),
        ("start", "<i4"),
        ("end", "<i4"),
        ("k", "<i4"),
        ("n", "<i4"),
        ("iters", "<i4"),
        ("metrics", "<f8", (4,)),
    ]
)
assert RECORD_DTYPE.itemsize == 64


class VspOpts(ctypes.Structure):
    _fields_ = [
        ("fit_start", c_int32),
        ("fit_end", c_int32),
        ("hill_k", c_int32),
        ("want_sv", c_int32),
        ("refine", c_int32),
        ("dist_k", c_int32),
        ("clauset", c_int32),
        ("reserved", c_int32 * 1),
    ]

    @classmethod
    def make(cls, fit_range=None, hill_k=None, want_sv=True, refine=None, dist_k=0, clauset=False):
        o = cls()
        o.fit_start, o.fit_end = (-1, -1) if fit_range is None else (int(fit_range[0]), int(fit_range[1]))
        o.hill_k = -1 if hill_k is None else int(hill_k)
        o.want_sv = 1 if want_sv else 0
        o.refine = -1 if refine is None else int(bool(refine))
        o.dist_k = max(0, int(dist_k or 0))
        o.clauset = 1 if clauset else 0
        return o


def aux_stride(dist_k: int, clauset: bool) -> int:
    """VSP_AUX_STRIDE of include/vspectra.h: doubles per matrix of the auxiliary output."""
    return 4 * max(0, int(dist_k or 0)) + (8 if clauset else 0)


class NativeError(RuntimeError):
    pass


TORCH_EXT_PATH = LIB_PATH.parent / "vspectra_torch.so"
_torch_ops = None


def load_torch_ext():
    """Load lib/vspectra_torch.so (csrc/torch_ext.cpp) and return `torch.ops.vision_spectra_b200`: the PyTorch
    extension layer over the C-ABI (`analyze_batch`: tensor list in, ATen-owned records / singular values out, on the
    current stream under a device guard).  Raises if it has not been built; there is no fallback op."""
    global _torch_ops
    if _torch_ops is not None:
        return _torch_ops
    import torch

    load()  # the C-ABI library first, so that the extension binds to the same copy (VSP_LIB override included)
    if not TORCH_EXT_PATH.exists():
        raise NativeError(f"{TORCH_EXT_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'`.")
    torch.ops.load_library(str(TORCH_EXT_PATH))
    _torch_ops = torch.ops.vision_spectra_b200
    return _torch_ops


_lib = None


def load() -> ctypes.CDLL:
    """Load lib/libvspectra.so and declare every symbol of include/vspectra.h."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise NativeError(
            f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). There is no CPU fallback for the spectral path."
        )
    lib = ctypes.CDLL(str(LIB_PATH))
    P32, P64 = POINTER(c_int32), POINTER(c_int64)
    sig = {
        "vsp_version": (c_int32, []),
        "vsp_error_string": (c_char_p, [c_int32]),
        "vsp_last_cuda_error": (c_char_p, []),
        "vsp_workspace_bytes": (c_int64, [c_int32, P32, P32]),
        "vsp_sv_offsets": (c_int32, [c_int32, P32, P32, P64]),
        "vsp_plan_create": (c_int32, [c_int32, P32, P32, P64, c_int32, POINTER(VspOpts), POINTER(c_void_p)]),
        "vsp_plan_workspace_bytes": (c_int64, [c_void_p]),
        "vsp_plan_sv_count": (c_int64, [c_void_p]),
        "vsp_plan_destroy": (None, [c_void_p]),
        "vsp_plan_execute": (c_int32, [c_void_p, POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
        "vsp_plan_execute_dist": (c_int32, [c_void_p, POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_void_p, c_int64, c_void_p]),
        "vsp_plan_execute_profiled": (
            c_int32,
            [c_void_p, POINTER(c_void_p), c_void_p, c_void_p, c_void_p, c_int64, c_void_p, POINTER(ctypes.c_float)],
        ),
        "vsp_plan_debug_gram": (c_int32, [c_void_p, POINTER(c_void_p), c_void_p, c_void_p, c_int64, c_void_p]),
        "vsp_analyze_batch": (
            c_int32,
            [POINTER(c_void_p), P32, P32, P64, c_int32, c_int32, POINTER(VspOpts), c_void_p, c_void_p, c_void_p, c_int64, c_void_p],
        ),
        "vsp_analyze_batch_host": (
            c_int32,
            [POINTER(c_void_p), P32, P32, P64, c_int32, c_int32, POINTER(VspOpts), POINTER(c_double), c_void_p, c_int32],
        ),
        "vsp_dgemm_batched": (
            c_int32,
            [c_int32, c_int32, c_int32, c_int32, c_double, c_void_p, c_int64, c_int32, c_void_p, c_int64, c_int32, c_double, c_double,
             c_void_p, c_int64, c_void_p],
        ),
        "vsp_kernel_launch_count": (c_int64, []),
        "vsp_reset_kernel_launch_count": (None, []),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)  # AttributeError here = header and library disagree
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


EXPORTED_SYMBOLS = (
    "vsp_version",
    "vsp_error_string",
    "vsp_last_cuda_error",
    "vsp_workspace_bytes",
    "vsp_sv_offsets",
    "vsp_plan_create",
    "vsp_plan_workspace_bytes",
    "vsp_plan_sv_count",
    "vsp_plan_destroy",
    "vsp_plan_execute",
    "vsp_plan_execute_dist",
    "vsp_dgemm_batched",
    "vsp_plan_execute_profiled",
    "vsp_plan_debug_gram",
    "vsp_analyze_batch",
    "vsp_analyze_batch_host",
    "vsp_kernel_launch_count",
    "vsp_reset_kernel_launch_count",
)


def check(rc: int, what: str = "vspectra") -> None:
    if rc >= 0:
        return
    lib = load()
    msg = lib.vsp_error_string(rc).decode()
    detail = lib.vsp_last_cuda_error().decode()
    raise NativeError(f"{what} failed: {msg} ({rc})" + (f" [{detail}]" if detail and rc == -4 else ""))


def i32(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int32)


def i64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.int64)


def p32(a: np.ndarray):
    return a.ctypes.data_as(POINTER(c_int32))


def p64(a: np.ndarray):
    return a.ctypes.data_as(POINTER(c_int64))
