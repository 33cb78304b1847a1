"""Small mixed batch through every kernel of the n <= 200 path (and one n = 300 matrix for the round-1 kernels), for
compute-sanitizer runs:  compute-sanitizer --tool memcheck python scripts/sanitize_case.py"""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vision_spectra_b200 as pkg

rng = np.random.default_rng(0)
shapes = [(8, 8), (9, 33), (33, 70), (64, 64), (100, 100), (137, 150), (192, 192), (768, 192), (200, 200), (300, 310)]
mats = [torch.from_numpy((rng.standard_normal(s) * 0.02).astype(np.float32)).cuda() for s in shapes]
u = np.linalg.qr(rng.standard_normal((64, 64)))[0]
v = np.linalg.qr(rng.standard_normal((64, 64)))[0]
mats.append(torch.from_numpy(((u * np.logspace(0, -7, 64)) @ v.T).astype(np.float32)).cuda())  # ill-conditioned: re-solve
eng = pkg.SpectraEngine(torch.device("cuda", 0))
d, c = [], []
metrics, svs, rec = eng.analyze(mats, dist_k=16, dist_out=d, clauset_out=c)
torch.cuda.synchronize()
print("status", [int(r["status"]) for r in rec], "alpha", [round(m["alpha_exponent"], 4) for m in metrics][:4])
from vision_spectra_b200.metrics.tail_truncation import truncate_weight_matrix
from vision_spectra_b200.metrics.gradient_alignment import compute_rank_reducing_gradient
t, info = truncate_weight_matrix(mats[4], 0.9)
p = compute_rank_reducing_gradient(mats[2])
torch.cuda.synchronize()
print("lowrank ok", info, tuple(p.shape))
