"""Recipe for oracle/_ref: the reference's OWN implementation of the hot path, placed where it can travel to the GPU box.

The reference is pure Python (no native code to compile): `vision_spectra/metrics/spectral.py` only imports NumPy and
SciPy at module level, so the file itself is the "built" artefact.  This script copies it, unmodified, from
/root/reference into oracle/_ref/ (git-ignored, NOT gpurun-ignored), together with a checksum file.  `bench.py --impl
reference` then times `get_spectral_metrics` of that very file plus the fifth `scipy.linalg.svd` the driver performs
(experiments/run_spectral_analysis.py:331-334) and reports `kind: "reference"`; without oracle/_ref it falls back to the
oracle port (`kind: "port"`).  tests/test_oracle.py checks the port against outputs of the same file (tests/golden).

    python oracle/build_ref.py            # __graft_entry__.build() runs this when /root/reference exists
"""
from __future__ import annotations

import hashlib
import shutil
import sys
from pathlib import Path

REF_FILE = Path("/root/reference/vision_spectra/metrics/spectral.py")
OUT_DIR = Path(__file__).resolve().parent / "_ref"


def build_ref(verbose: bool = True) -> Path | None:
    if not REF_FILE.exists():
        if verbose:
            print(f"[oracle/_ref] {REF_FILE} not present (GPU box): keeping what is already in {OUT_DIR}")
        return OUT_DIR / "ref_spectral.py" if (OUT_DIR / "ref_spectral.py").exists() else None
    OUT_DIR.mkdir(parents=True, exist_ok=True)
    dst = OUT_DIR / "ref_spectral.py"
    shutil.copyfile(REF_FILE, dst)
    digest = hashlib.sha256(dst.read_bytes()).hexdigest()
    (OUT_DIR / "SOURCE.txt").write_text(f"{REF_FILE}\nsha256 {digest}\ncopied unmodified by oracle/build_ref.py\n")
    if verbose:
        print(f"[oracle/_ref] {dst} (sha256 {digest[:16]}...)")
    return dst


def load_ref():
    """Import oracle/_ref/ref_spectral.py as a module, or None if it has not been built."""
    import importlib.util

    src = OUT_DIR / "ref_spectral.py"
    if not src.exists():
        return None
    spec = importlib.util.spec_from_file_location("ref_spectral", src)
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_spectral"] = mod
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    build_ref()
