"""Development: run a batch of n x n matrices through the CUDA path (phase-timing builds print cycle counters)."""
import sys, os, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_spectra_b200.engine import analyze_matrices
n = int(sys.argv[1]) if len(sys.argv) > 1 else 192
count = int(sys.argv[2]) if len(sys.argv) > 2 else 1200
g = torch.Generator(device="cuda").manual_seed(0)
mats = [torch.randn(n, n, generator=g, device="cuda") * 0.02 for _ in range(count)]
for _ in range(2):
    torch.cuda.synchronize(); t = time.time()
    analyze_matrices(mats)
    torch.cuda.synchronize(); print("wall ms", (time.time() - t) * 1e3)
