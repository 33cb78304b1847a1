"""e2e-only probe: times SweepRunner.run_host on the Scenario-A sweep (pinned host arenas); CHUNK / LANES from the environment."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vision_spectra_b200.engine import SpectraEngine
from vision_spectra_b200.sweep import CheckpointLayout, SweepRunner

dev = torch.device("cuda:0")
lay = CheckpointLayout.vit(192, 6)
nck = 93
host_block = torch.empty(nck * lay.arena_elems, dtype=torch.float32).pin_memory()
g = torch.Generator().manual_seed(1)
host_block.normal_(generator=g).mul_(0.02)
arenas = [host_block[i * lay.arena_elems:(i + 1) * lay.arena_elems] for i in range(nck)]
chunk = int(os.environ.get("CHUNK", "8"))
lanes = int(os.environ.get("LANES", "6"))
runner = SweepRunner(SpectraEngine(dev), lay, ckpts_per_chunk=chunk, lanes=lanes)
for _ in range(3):
    runner.run_host(arenas)
torch.cuda.synchronize()
ts = []
for _ in range(8):
    t0 = time.perf_counter()
    runner.run_host(arenas)
    ts.append(time.perf_counter() - t0)
ts = np.array(ts) * 1e3
print(f"chunk {chunk} lanes {lanes}: ms median {np.median(ts):.2f} min {ts.min():.2f}  -> {nck * lay.matrices / np.median(ts) * 1e3:.0f} matrices/s; floor at 55 GB/s {host_block.numel() * 4 / 55e9 * 1e3:.2f} ms")
