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
TORCH_EXT = PKG / "lib" / "vspectra_torch.so"

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
    build_torch_ext(force=True)
    return LIB


def build_torch_ext(force: bool = False) -> Path:
    """The PyTorch extension layer (csrc/torch_ext.cpp: torch.ops.vision_spectra_b200.analyze_batch), a host-only C++
    file over the C-ABI: g++ against the torch headers, linked to lib/libvspectra.so (rpath $ORIGIN), in-tree."""
    import torch
    from torch.utils import cpp_extension as ce

    src = CSRC / "torch_ext.cpp"
    if not force and TORCH_EXT.exists() and TORCH_EXT.stat().st_mtime > max(src.stat().st_mtime, LIB.stat().st_mtime):
        return TORCH_EXT
    cuda_home = Path(_nvcc()).resolve().parent.parent
    tlib = Path(torch.__file__).resolve().parent / "lib"
    cmd = ["g++", "-O2", "-std=c++17", "-shared", "-fPIC", f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}",
           "-DTORCH_EXTENSION_NAME=vspectra_torch"]
    cmd += [f"-I{p}" for p in ce.include_paths()] + [f"-I{cuda_home / 'include'}"]
    cmd += [str(src), "-o", str(TORCH_EXT), f"-L{tlib}", "-ltorch", "-ltorch_cpu", "-ltorch_cuda", "-lc10", "-lc10_cuda",
            f"-L{LIB.parent}", "-lvspectra", f"-L{cuda_home / 'lib64'}", "-lcudart", "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{tlib}"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed (torch_ext.cpp):\n" + res.stdout + res.stderr[-4000:])
    return TORCH_EXT


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
