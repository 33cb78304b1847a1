"""Device-resident throughput of the other BASELINE.json configs (the bench line itself is Scenario A):
E (32d/1L), C (96d/3L), A (192d/6L), the six-scenario sweep (A, D, E x 31 + B, C, F x 51 checkpoints x 3 seeds, as
shape classes 192/96/32) and a ViT-Base (768d/12L) sample.  Prints one JSON line per config."""
import json, os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import vision_spectra_b200 as pkg
from vision_spectra_b200.sweep import CheckpointLayout, SweepRunner

dev = torch.device("cuda", 0)
eng = pkg.SpectraEngine(dev)


def run(name, parts, reps=3):
    """parts: list of (embed_dim, depth, n_checkpoints)"""
    runners = []
    for d, L, n in parts:
        lay = CheckpointLayout.vit(d, L)
        g = torch.Generator(device=dev).manual_seed(d * 1000 + L)
        arenas = [torch.randn(lay.arena_elems, generator=g, device=dev) * 0.02 for _ in range(n)]
        runners.append((SweepRunner(eng, lay), arenas, lay))
    mats = sum(len(a) * lay.matrices for _, a, lay in runners)
    for r, a, _ in runners:
        r.run_device(a)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for r, a, _ in runners:
            res = r.run_device(a)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    rec = res.records_host()
    print(json.dumps({"config": name, "matrices": mats, "ms": round(ms, 3), "matrices_per_s": round(mats / ms * 1e3),
                      "clean": bool(((rec["status"] == 0) | (rec["status"] == 96)).all())}))


run("E: ViT 32d/1L x 93 checkpoints", [(32, 1, 93)])
run("C: ViT 96d/3L x 153 checkpoints", [(96, 3, 153)])
run("A: ViT 192d/6L x 93 checkpoints", [(192, 6, 93)])
run("six-scenario sweep (14 760 matrices)", [(192, 6, 93), (96, 3, 93), (32, 1, 93), (192, 6, 153), (96, 3, 153), (32, 1, 153)])
run("Base: ViT 768d/12L x 2 checkpoints", [(768, 12, 2)], reps=1)
