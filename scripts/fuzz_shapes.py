"""One-off sweep: random shapes through the device path against NumPy's SVD (singular values, element-wise gate)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vision_spectra_b200 as pkg

rng = np.random.default_rng(int(os.environ.get("SEED", "3")))
eng = pkg.SpectraEngine(torch.device("cuda", 0))
host = []
for _ in range(int(os.environ.get("COUNT", "160"))):
    n = int(rng.integers(8, 330))
    K = int(rng.choice([n, n + int(rng.integers(0, 40)), 2 * n, 4 * n]))
    w = (rng.standard_normal((n, K)) * 0.02).astype(np.float32)
    host.append(w if rng.random() < 0.5 else np.ascontiguousarray(w.T))
metrics, svs, rec = eng.analyze([torch.from_numpy(w).cuda() for w in host])
worst = 0.0
for w, s, r in zip(host, svs, rec):
    ref = np.linalg.svd(w.astype(np.float64), compute_uv=False)
    err = np.max(np.abs(np.asarray(s) - ref) / ref)
    worst = max(worst, err)
    assert int(r["status"]) in (0, 96), (w.shape, int(r["status"]))
    assert err < 1e-5, (w.shape, err, int(r["status"]))
print("fuzz ok:", len(host), "matrices, worst relative singular-value error", worst)
