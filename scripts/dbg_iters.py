"""Development: bisection iteration counts (max per matrix) reported in the records."""
import sys, os
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vision_spectra_b200 as pkg
eng = pkg.SpectraEngine(torch.device("cuda", 0))
g = torch.Generator().manual_seed(0)
mats = [torch.randn(192, 192, generator=g) * 0.02 for _ in range(40)] + [torch.randn(768, 192, generator=g) * 0.02 for _ in range(20)]
metrics, svs, rec = eng.analyze([m.cuda() for m in mats])
it = np.array([int(r["iters"]) for r in rec])
print("square 192: iters min/mean/max", it[:40].min(), it[:40].mean(), it[:40].max())
print("768x192   : iters min/mean/max", it[40:].min(), it[40:].mean(), it[40:].max())
