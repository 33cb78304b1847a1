"""Build the CUDA library in-tree with nvcc for sm_100a.

    python -m vision_spectra_b200.build            # or __graft_entry__.build()

The output (`lib/libvspectra.so`) is git-ignored but travels to the GPU box with
the repository snapshot.  nvcc cross-compiles without a GPU.
"""

from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "lib" / "libvspectra.so"

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-O3",
    "-std=c++17",
    "--expt-extended-lambda",
    "-shared",
    "-Xcompiler",
    "-fPIC",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; the CUDA path cannot be built (there is no CPU fallback)")


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def is_stale() -> bool:
    if not LIB.exists():
        return True
    t = LIB.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "vspectra.h"]
    return any(d.stat().st_mtime > t for d in deps)


def build_native(force: bool = False, verbose: bool = False) -> Path:
    if not force and not is_stale():
        return LIB
    LIB.parent.mkdir(parents=True, exist_ok=True)
    extra = os.environ.get("VSP_EXTRA_NVCC_FLAGS", "").split()  # development switches, e.g. -DVSP_PHASE_TIMING
    cmd = [_nvcc(), *NVCC_FLAGS, *extra, "-o", str(LIB), *[str(s) for s in sources()]]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    if verbose:
        print(res.stderr)
    build_micro()
    return LIB


def build_micro() -> Path:
    """Peak microbenchmarks bench.py runs on the GPU box for the roofline denominators MEASURED_PEAKS.json does not
    have: FP64 tensor cores (DMMA.8x8x4) and int8 tcgen05.mma."""
    exe = LIB.parent / "dmma_bench"
    for name in ("dmma_bench", "i8_umma_bench"):
        src = PKG.parent / "scripts" / "micro" / f"{name}.cu"
        if src.exists():
            res = subprocess.run([_nvcc(), "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-o", str(LIB.parent / name), str(src)],
                                 capture_output=True, text=True)
            if res.returncode != 0:
                raise RuntimeError(f"nvcc failed ({name}):\n" + res.stdout + res.stderr)
    return exe


if __name__ == "__main__":
    print(build_native(force=True, verbose=True))
