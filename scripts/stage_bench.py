"""Development: per-stage (and, with VSP_KERNEL_TIMING=1, per-launch) times of one Scenario-A sweep, device-resident.
    python scripts/stage_bench.py [embed_dim depth ckpts]"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vision_spectra_b200 as pkg
from vision_spectra_b200.sweep import CheckpointLayout, SweepRunner

d = int(sys.argv[1]) if len(sys.argv) > 1 else 192
depth = int(sys.argv[2]) if len(sys.argv) > 2 else 6
nck = int(sys.argv[3]) if len(sys.argv) > 3 else 93
dev = torch.device("cuda", 0)
lay = CheckpointLayout.vit(d, depth)
eng = pkg.SpectraEngine(dev)
r = SweepRunner(eng, lay)
g = torch.Generator(device=dev).manual_seed(1)
arenas = [torch.randn(lay.arena_elems, generator=g, device=dev) * 0.02 for _ in range(nck)]
for _ in range(3):
    res = r.run_device(arenas)
torch.cuda.synchronize()
acc = [0.0, 0.0, 0.0]
for _ in range(3):
    sm = []
    res = r.run_device(arenas, stage_ms=sm)
    acc = [a + b for a, b in zip(acc, sm)]
rec = res.records_host()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    r.run_device(arenas)
e1.record()
torch.cuda.synchronize()
import numpy as np
print("stage ms gram/reduce/bisect:", [round(a / 3, 3) for a in acc], "step ms", round(e0.elapsed_time(e1) / 5, 3),
      "refined", int((rec["status"] & 32).astype(bool).sum()), "bad", int(((rec["status"] != 0) & (rec["status"] != 96)).sum()),
      "chk", float(np.nansum(rec["metrics"])))
