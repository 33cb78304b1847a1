"""Development: one ViT-Base checkpoint (72 matrices, n = 768) through the device path."""
import sys, os, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vision_spectra_b200 as pkg
from vision_spectra_b200.sweep import CheckpointLayout, SweepRunner
dev = torch.device("cuda", 0)
eng = pkg.SpectraEngine(dev)
lay = CheckpointLayout.vit(768, 12)
g = torch.Generator(device=dev).manual_seed(1)
arenas = [torch.randn(lay.arena_elems, generator=g, device=dev) * 0.02 for _ in range(2)]
r = SweepRunner(eng, lay)
for _ in range(2):
    torch.cuda.synchronize(); t = time.time(); r.run_device(arenas); torch.cuda.synchronize(); print("ms", (time.time() - t) * 1e3)
