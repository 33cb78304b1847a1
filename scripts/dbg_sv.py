"""Development check: singular values from the CUDA path vs numpy SVD for a ladder of shapes."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_spectra_b200.engine import analyze_matrices

shapes = [(3, 3), (5, 7), (6, 6), (8, 8), (13, 20), (32, 32), (33, 33), (40, 64), (64, 64), (65, 130), (96, 96), (100, 100),
          (128, 128), (160, 200), (192, 192), (768, 192), (192, 768), (200, 200), (256, 256)]
if len(sys.argv) > 1:
    shapes = [tuple(int(x) for x in a.split("x")) for a in sys.argv[1:]]
g = torch.Generator().manual_seed(0)
mats = [torch.randn(r, c, generator=g) * 0.02 for r, c in shapes]
mets, svs = analyze_matrices([m.cuda() for m in mats])
bad = 0
for (r, c), m, sv in zip(shapes, mats, svs):
    ref = np.linalg.svd(m.double().numpy(), compute_uv=False)
    sv = np.asarray(sv)
    if sv.shape != ref.shape or not np.all(np.isfinite(sv)):
        print(r, c, "BAD shape/finite", sv.shape, np.isfinite(sv).sum()); bad += 1; continue
    rel = np.abs(sv - ref) / ref
    print(f"{r}x{c}: max rel {rel.max():.3e}  max abs/smax {np.abs(sv-ref).max()/ref[0]:.3e}")
    if rel.max() > 1e-6: bad += 1
print("BAD" if bad else "ALL OK", bad)
