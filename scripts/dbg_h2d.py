"""Development: raw pinned H2D bandwidth on this box, and the e2e sweep at several chunk / lane settings."""
import sys, os, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
dev = torch.device("cuda", 0)
h = torch.empty(1 << 30, dtype=torch.uint8).pin_memory()
d = torch.empty(1 << 30, dtype=torch.uint8, device=dev)
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("H2D 1 GiB pinned:", 5 * (1 << 30) / (e0.elapsed_time(e1) / 1e3) / 1e9, "GB/s")
# many 10.6 MB copies
hs = [torch.empty(10616832, dtype=torch.uint8).pin_memory() for _ in range(93)]
e0.record()
for _ in range(3):
    for i, x in enumerate(hs): d[i * 10616832:(i + 1) * 10616832].copy_(x, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("H2D 93 x 10.6 MB pinned:", 3 * 93 * 10616832 / (e0.elapsed_time(e1) / 1e3) / 1e9, "GB/s")
import vision_spectra_b200 as pkg
from vision_spectra_b200.sweep import CheckpointLayout, SweepRunner
lay = CheckpointLayout.vit(192, 6)
block = (torch.randn(93 * lay.arena_elems) * 0.02).pin_memory()
host = [block[i * lay.arena_elems:(i + 1) * lay.arena_elems] for i in range(93)]
for chunk, lanes in [(8, 6), (8, 6), (4, 8), (16, 6)]:
    eng = pkg.SpectraEngine(dev)
    r = SweepRunner(eng, lay, ckpts_per_chunk=chunk, lanes=lanes)
    r.run_host(host); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(6): r.run_host(host)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t) / 6
    print(f"chunk {chunk} lanes {lanes}: {93*36/dt:.0f} matrices/s  ({93*lay.bytes/dt/1e9:.1f} GB/s H2D equivalent)")
    del r, eng
