#!/usr/bin/env python
"""bench.py -- weight matrices spectrally analysed per second (BASELINE.json metric).

    python bench.py --gpus 1 --steps 5 --warmup 3                 # this repo's CUDA path
    python bench.py --impl reference --gpus 1 --steps 2 --warmup 1 # reference CPU algorithm
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
        --master-port P bench.py --gpus N --steps K --warmup W   # one rank per GPU

Workload (config.workload): BASELINE.json configs[1], "Scenario A: ViT-Tiny 192d/6L":
36 fp32 matrices per checkpoint (24x 192x192 as q/k/v row-blocks of a fused qkv buffer
+ proj, 6x 768x192, 6x 192x768), 31 epochs x 3 seeds = 93 checkpoints = 3348 matrices =
987 MB per GPU per step (larger than the 126 MB L2, so every step re-reads HBM).
One step = one pass of the hot path over that batch.  Weak scaling: every rank analyses
its own 93 checkpoints, the only exchange is one NCCL gather of the 64-byte records.

Printed keys: see the contract in the task description; `value` = device-resident
throughput, `e2e` = same metric from pinned HOST buffers through the public API
(`SweepRunner.run_host`: H2D copies + kernels + D2H of records and singular values in the
timed region), `roofline` = the dominant kernel against its bound, `cpu_baseline` = the
reference's algorithm (oracle port: 5 SciPy SVDs per matrix) timed on this box's cores.
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

SCENARIO = {"name": "Scenario A: ViT-Tiny 192d/6L", "embed_dim": 192, "depth": 6, "epochs": 31, "seeds": (42, 142, 242)}
FP64_PEAK_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12  # 148 SM x 64 FP64 FMA/clk x 2 flop x 1.965 GHz


# --------------------------------------------------------------------------- utils
def measure_fp64_peak():
    """FP64 peak of this GPU, measured live with the DMMA.8x8x4 microbenchmark that `build()` compiles
    (scripts/micro/dmma_bench.cu; MEASURED_PEAKS.json has no FP64 figure).  Falls back to the derived number."""
    import subprocess

    exe = ROOT / "vision-spectra_b200" / "lib" / "dmma_bench"
    try:
        out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60).stdout
        vals = [float(ln.split(":")[1].split()[0]) for ln in out.splitlines() if ln.startswith("dmma m8n8k4")]
        if vals:
            return max(vals), "measured live: FP64 tensor-core (DMMA.8x8x4) microbenchmark, scripts/micro/dmma_bench.cu"
    except Exception:
        pass
    return FP64_PEAK_TFLOPS, "derived: 148 SM x 64 FP64 FMA/clk x 1.965 GHz (no FP64 figure in MEASURED_PEAKS.json)"


def measure_i8_peak():
    """int8 tensor-core peak of this GPU (tcgen05.mma kind::i8), measured live with scripts/micro/i8_umma_bench.cu:
    (TOPS at the production shape M=128 N=64 K=32, TOPS at N=256, source).  Falls back to 2x the measured bf16
    figure of MEASURED_PEAKS.json, said so in the source string."""
    import subprocess

    exe = ROOT / "vision-spectra_b200" / "lib" / "i8_umma_bench"
    try:
        out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=60).stdout
        vals = {ln.split(":")[0].strip(): float(ln.split(":")[1].split()[0]) for ln in out.splitlines() if ln.startswith("i8 umma")}
        if "i8 umma m128n64k32" in vals and "i8 umma m128n256k32" in vals:
            return vals["i8 umma m128n64k32"], vals["i8 umma m128n256k32"], "measured live: tcgen05.mma kind::i8 issue-loop microbenchmark, scripts/micro/i8_umma_bench.cu"
    except Exception:
        pass
    d = load_peaks()
    return 2.0 * d["bf16_tflops"], 2.0 * d["bf16_tflops"], "derived: 2 x bf16 burst of " + d["source"] + " (microbenchmark did not run)"


def load_traffic():
    """DRAM bytes per launch of the dominant kernel from the committed ncu capture (latest profiles/r*_traffic.json)."""
    try:
        cand = sorted((ROOT / "profiles").glob("r*_traffic.json"))
        t = json.loads(cand[-1].read_text())
        return t["dram_bytes_read"] + t["dram_bytes_write"], f'{t["kernel"]}: {t["source"]}' 
    except Exception:
        return None, None


def load_peaks() -> dict:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        d["source"] = "measured (MEASURED_PEAKS.json)"
        return d
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback (B200_PROFILING.md)"}


class ClockSampler:
    """Samples SM clock and throttle reasons of one GPU during the timed region (NVML)."""

    REASONS = {
        0x0000000000000004: "sw_power_cap",
        0x0000000000000008: "hw_slowdown",
        0x0000000000000020: "sw_thermal_slowdown",
        0x0000000000000040: "hw_thermal_slowdown",
        0x0000000000000080: "hw_power_brake_slowdown",
    }

    def __init__(self, index: int):
        self.index, self.samples, self.reasons, self.max_mhz = index, [], set(), None
        self._stop = threading.Event()
        self._thread = None
        try:
            import pynvml

            pynvml.nvmlInit()
            self._nv = pynvml
            self._h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self._nv = None

    def _run(self):
        nv = self._nv
        while not self._stop.is_set():
            try:
                self.samples.append(float(nv.nvmlDeviceGetClockInfo(self._h, nv.NVML_CLOCK_SM)))
                mask = int(nv.nvmlDeviceGetCurrentClocksEventReasons(self._h))
                for bit, name in self.REASONS.items():
                    if mask & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.004)  # the timed region of the default run is ~50 ms

    def __enter__(self):
        if self._nv is not None:
            self._thread = threading.Thread(target=self._run, daemon=True)
            self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thread is not None:
            self._thread.join(timeout=2)

    def summary(self) -> dict:
        import statistics

        return {
            "sm_mhz": statistics.median(self.samples) if self.samples else None,
            "sm_max_mhz": self.max_mhz,
            "reasons": sorted(self.reasons),
            "samples": len(self.samples),
        }


def physical_gpu_index(local: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except Exception:
            return local
    return local


# ------------------------------------------------------------------- CPU reference
def _cpu_worker_init():
    """One BLAS thread per worker process.  Both OpenBLAS copies (NumPy's and SciPy's) must
    be loaded before the limit is applied, or SciPy's keeps its default thread count and
    the pool oversubscribes the cores."""
    global _blas_limit
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    os.environ["OMP_NUM_THREADS"] = "1"
    import numpy  # noqa: F401
    import scipy.linalg  # noqa: F401
    import scipy.stats  # noqa: F401
    from threadpoolctl import threadpool_limits

    _blas_limit = threadpool_limits(1)


_ref_mod = None


def reference_kind() -> str:
    """"reference": oracle/_ref holds the reference's own metrics/spectral.py (oracle/build_ref.py); "port": it does
    not, the oracle restatement is timed instead."""
    return "reference" if (ROOT / "oracle" / "_ref" / "ref_spectral.py").exists() else "port"


def _cpu_worker(w):
    """One matrix the way the reference's driver treats it (experiments/run_spectral_analysis.py:323-336):
    get_spectral_metrics (4 SVDs) + the fifth SVD that stores the singular values."""
    global _ref_mod
    sys.path.insert(0, str(ROOT / "oracle"))
    if reference_kind() == "reference":
        if _ref_mod is None:
            import build_ref

            _ref_mod = build_ref.load_ref()
        import numpy as np
        from scipy.linalg import svd

        m = _ref_mod.get_spectral_metrics(w)
        svd(np.asarray(w, dtype=np.float64), compute_uv=False)
        return m["stable_rank"]
    import spectral_oracle as orc

    out = orc.reference_cost_metrics(w)  # 4 metric SVDs + the driver's 5th SVD
    return out["metrics"]["stable_rank"]


def host_checkpoints(n_ckpt: int, seed0: int = 42):
    """Synthetic Scenario-A checkpoints on the host (NumPy), the reference arm's input."""
    import numpy as np

    d, depth = SCENARIO["embed_dim"], SCENARIO["depth"]
    out = []
    for c in range(n_ckpt):
        rng = np.random.default_rng(seed0 * 1_000_003 + c)
        mats = []
        for _ in range(depth):
            qkv = (rng.standard_normal((3 * d, d)) * 0.02).astype(np.float32)
            mats += [qkv[:d], qkv[d : 2 * d], qkv[2 * d :]]
            mats += [(rng.standard_normal(s) * 0.02).astype(np.float32) for s in ((d, d), (4 * d, d), (d, 4 * d))]
        out.append(mats)
    return out


def time_cpu_reference(n_ckpt: int, cores: int, pool=None) -> tuple[float, int]:
    """Seconds to analyse n_ckpt checkpoints with the reference's algorithm on `cores`
    processes x 1 BLAS thread (BASELINE.md "pool" mode).  Returns (seconds, matrices)."""
    mats = [w for ck in host_checkpoints(n_ckpt) for w in ck]
    own = pool is None
    if own:
        import multiprocessing as mp

        pool = mp.get_context("fork").Pool(cores, initializer=_cpu_worker_init)
        pool.map(_cpu_worker, mats[: min(len(mats), cores)])  # warm the workers
    t0 = time.perf_counter()
    pool.map(_cpu_worker, mats, chunksize=1)
    dt = time.perf_counter() - t0
    if own:
        pool.close()
        pool.join()
    return dt, len(mats)


def host_cores() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def run_reference_arm(args) -> None:
    """`--impl reference`: the reference's own CPU implementation of this path on all host cores: the reference's
    metrics/spectral.py itself when oracle/_ref has been built (oracle/build_ref.py, `kind: "reference"`), else the
    oracle port (same SciPy/LAPACK calls, `kind: "port"`); 5 SVDs per matrix either way."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import multiprocessing as mp

    cores = host_cores()
    # bounded sample, but at least ~4 matrices per core so every core is kept busy
    n_ckpt = max(1, args.ref_ckpts, -(-cores * 4 // 36))
    pool = mp.get_context("fork").Pool(cores, initializer=_cpu_worker_init)
    mats = [w for ck in host_checkpoints(n_ckpt) for w in ck]
    times = []
    for step in range(args.warmup + args.steps):
        t0 = time.perf_counter()
        pool.map(_cpu_worker, mats, chunksize=1)
        dt = time.perf_counter() - t0
        if step >= args.warmup:
            times.append(dt)
    pool.close()
    pool.join()
    total = sum(times)
    value = len(mats) * len(times) / total
    sample = f"{n_ckpt} Scenario-A checkpoints ({len(mats)} matrices) per step, {cores} processes x 1 BLAS thread"
    line = {
        "impl": "reference",
        "metric": "weight matrices spectrally analysed/sec (ViT-Tiny Q/K/V/MLP)",
        "value": value,
        "unit": "matrices/s",
        "n_gpus": args.gpus,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times),
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": SCENARIO["name"], "matrices_per_step": len(mats), "sample": sample},
        "cpu_baseline": {"value": value, "unit": "matrices/s", "cores": cores, "kind": reference_kind(), "sample": sample},
        "e2e": {"value": value, "unit": "matrices/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)



# ------------------------------------------------------------- extras (outside the headline timed region)
SCENARIOS = {  # reference table experiments/run_spectral_analysis.py:145-236: (embed_dim, depth, epochs incl. epoch 0)
    "A": (192, 6, 31), "B": (192, 6, 51), "C": (96, 3, 51), "D": (96, 3, 31), "E": (32, 1, 31), "F": (32, 1, 51),
}
SEEDS = (42, 142, 242)  # run_spectral_analysis.py:706


def _device_time(fn, sync_all, steps: int, warmup: int = 2, collective: bool = True) -> float:
    """ms per call of fn(), CUDA events on the current stream, barrier + synchronize on both sides; max over ranks when
    every rank takes part (`collective`), this rank's own time for the rank-0-only entries."""
    import torch
    import torch.distributed as dist

    for _ in range(warmup):
        fn()
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fn()
    e1.record()
    sync_all()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda", dtype=torch.float64)
    if collective and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    return float(ms.item())


def extra_six_scenario_sweep(eng, dev, world: int, rank: int, sync_all, steps: int = 3) -> dict:
    """BASELINE.json configs[3]: the full six-scenario x 3-seed x every-epoch sweep (14 760 matrices in three shape
    classes), STRONG-scaled: the 738 checkpoints are dealt to the ranks by longest-processing-time
    (sweep.partition_lpt on the per-checkpoint cost), every rank analyses its shard in ONE batched call and rank 0
    receives one ragged gather of 64-byte records."""
    import numpy as np
    import torch

    from vision_spectra_b200 import _native as nat
    from vision_spectra_b200.sweep import CheckpointLayout, gather_records_ragged, matrix_cost, partition_lpt

    layouts = {k: CheckpointLayout.vit(d, depth) for k, (d, depth, _) in SCENARIOS.items()}
    ckpts = [(k, seed, ep) for k, (_, _, epochs) in SCENARIOS.items() for seed in SEEDS for ep in range(epochs)]
    costs = [sum(matrix_cost(sl.rows, sl.cols) for sl in layouts[k].slots) for k, _, _ in ckpts]
    shard = partition_lpt(costs, world)[rank]
    arenas, ptrs, rows, cols = [], [], [], []
    for ci in shard:
        k, seed, ep = ckpts[ci]
        lay = layouts[k]
        g = torch.Generator(device=dev).manual_seed(seed * 1_000_003 + ep * 7 + ord(k))
        a = torch.randn(lay.arena_elems, generator=g, device=dev, dtype=torch.float32) * 0.02
        arenas.append(a)
        base = np.uint64(a.data_ptr())
        ptrs.append(base + np.array([4 * sl.offset for sl in lay.slots], dtype=np.uint64))
        rows.append(np.array([sl.rows for sl in lay.slots], np.int32))
        cols.append(np.array([sl.cols for sl in lay.slots], np.int32))
    ptrs, rows, cols = np.concatenate(ptrs), np.concatenate(rows), np.concatenate(cols)
    ld = cols.astype(np.int64)
    plan = eng.make_plan(rows, cols, ld, nat.VSP_F32, want_sv=True)
    state = {}

    def step():
        res = eng.analyze_raw(ptrs, rows, cols, ld, nat.VSP_F32, want_sv=True, plan=plan)
        state["res"] = res
        state["gathered"] = gather_records_ragged(res.records) if world > 1 else res.records

    ms = _device_time(step, sync_all, steps)
    rec = state["res"].records_host()
    ok = bool(((rec["status"] == 0) | (rec["status"] == 96)).all())
    total = sum(len(layouts[k].slots) for k, _, _ in ckpts)
    if rank == 0 and world > 1:
        assert state["gathered"].numel() == total * 64, "ragged gather lost records"
    del arenas
    return {"workload": "six scenarios x 3 seeds x every epoch (A,D,E 31 ckpts; B,C,F 51 ckpts)", "matrices": total,
            "checkpoints": len(ckpts), "scaling": "strong", "n_gpus": world, "ms": ms, "matrices_per_s": total / (ms / 1e3),
            "shard_matrices_rank0": int(len(ptrs)), "records_clean": ok,
            "partition": "sweep.partition_lpt over checkpoints; gather_records_ragged of 64-byte records"}


def extra_scenario(eng, dev, key: str, sync_all, steps: int = 3) -> dict:
    """One scenario's sweep (every epoch x 3 seeds), device-resident, on this rank (BASELINE.json configs[0] / [2])."""
    import torch

    from vision_spectra_b200.sweep import CheckpointLayout, SweepRunner

    d, depth, epochs = SCENARIOS[key]
    lay = CheckpointLayout.vit(d, depth)
    runner = SweepRunner(eng, lay)
    g = torch.Generator(device=dev).manual_seed(1000 + ord(key))
    arenas = [torch.randn(lay.arena_elems, generator=g, device=dev, dtype=torch.float32) * 0.02 for _ in range(epochs * len(SEEDS))]
    ms = _device_time(lambda: runner.run_device(arenas, want_sv=True), sync_all, steps, collective=False)
    n = len(arenas) * lay.matrices
    return {"workload": f"Scenario {key}: ViT {d}d/{depth}L, {epochs} epochs x {len(SEEDS)} seeds", "matrices": n, "ms": ms,
            "matrices_per_s": n / (ms / 1e3), "per_rank": True}


def extra_vit_base_stress(eng, dev, sync_all, ckpts: int = 16, chunk: int = 4) -> dict:
    """BASELINE.json configs[4]: ViT-Base 768d/12L random-init checkpoints (72 matrices each, up to 768x3072), analysed
    in chunks of `chunk` checkpoints (1.36 GB of inputs per chunk, regenerated per chunk: 10 000 checkpoints would be
    3.4 TB).  Chunks alternate between two lanes (engine + stream + arena set each) and a chunk's records are read one
    chunk later, so the re-solve tail of chunk i (19 ms on a few 16-CTA clusters) runs beside chunk i+1's reduction.
    Reports matrices/s of this rank and the stage times / FP64 roofline of the n = 768 reduction."""
    import torch

    from vision_spectra_b200.engine import SpectraEngine
    from vision_spectra_b200.sweep import CheckpointLayout, SweepRunner

    lay = CheckpointLayout.vit(768, 12)
    g = torch.Generator(device=dev).manual_seed(768)
    lanes = []
    for k in range(2):
        lane_eng = eng if k == 0 else SpectraEngine(dev)
        arenas = [torch.randn(lay.arena_elems, generator=g, device=dev, dtype=torch.float32) * 0.02 for _ in range(chunk)]
        lanes.append((SweepRunner(lane_eng, lay), torch.cuda.Stream(device=dev), arenas))
    for runner, stream, arenas in lanes:  # warm-up (plan, workspace)
        with torch.cuda.stream(stream):
            runner.run_device(arenas, want_sv=True)
    sync_all()
    main = torch.cuda.current_stream(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    refined = 0
    pending = None
    e0.record()
    for s in (l[1] for l in lanes):
        s.wait_stream(main)
    for ci, c in enumerate(range(0, ckpts, chunk)):
        runner, stream, arenas = lanes[ci % 2]
        with torch.cuda.stream(stream):
            for a in arenas:  # the next chunk's checkpoints (device-side generation is part of the chunk loop, not of the metric)
                a.normal_(0.0, 0.02, generator=g)
            res = runner.run_device(arenas, want_sv=True)
        if pending is not None:
            with torch.cuda.stream(pending[1]):
                refined += int((pending[0].records_host()["status"] == 96).sum())
        pending = (res, stream)
    with torch.cuda.stream(pending[1]):
        refined += int((pending[0].records_host()["status"] == 96).sum())
    for s in (l[1] for l in lanes):
        main.wait_stream(s)
    e1.record()
    sync_all()
    ms = e0.elapsed_time(e1)
    sm: list = []
    lanes[0][0].run_device(lanes[0][2], want_sv=True, stage_ms=sm)
    sync_all()
    n = ckpts * lay.matrices
    flops = chunk * lay.flops_tridiag()
    return {"workload": f"ViT-Base 768d/12L, {ckpts} checkpoints in chunks of {chunk}, two lanes", "matrices": n, "ms": ms,
            "matrices_per_s": n / (ms / 1e3), "refined": refined, "per_rank": True,
            "stage_ms_per_chunk": {"gram": sm[0], "reduction": sm[1], "bisect_refine": sm[2]},
            "reduction_tflops": flops / 1e12 / (sm[1] / 1e3), "note": "timed region includes regenerating each chunk's weights on the device"}


def extra_single_checkpoint_latency(dev, reps: int = 30) -> dict:
    """What a user of the drop-in sees: `extract_and_analyze_weights(model, device)` on a live ViT-Tiny-shaped module
    (192d / 6 blocks, timm's module names; run_spectral_analysis.py:297-345), wall clock per call, result dict on the host."""
    import statistics

    import torch
    import torch.nn as nn

    from vision_spectra_b200.experiments.run_spectral_analysis import extract_and_analyze_weights

    class Attn(nn.Module):
        def __init__(self, d):
            super().__init__()
            self.qkv, self.proj = nn.Linear(d, 3 * d), nn.Linear(d, d)

    class Mlp(nn.Module):
        def __init__(self, d):
            super().__init__()
            self.fc1, self.fc2 = nn.Linear(d, 4 * d), nn.Linear(4 * d, d)

    class Block(nn.Module):
        def __init__(self, d):
            super().__init__()
            self.attn, self.mlp = Attn(d), Mlp(d)

    class Vit(nn.Module):
        def __init__(self, d, depth):
            super().__init__()
            self.blocks = nn.Sequential(*[Block(d) for _ in range(depth)])

    torch.manual_seed(0)
    model = Vit(192, 6).to(dev)
    out = None
    for _ in range(3):
        out = extract_and_analyze_weights(model, dev)
    times = []
    for _ in range(reps):
        torch.cuda.synchronize(dev)
        t0 = time.perf_counter()
        out = extract_and_analyze_weights(model, dev)
        times.append((time.perf_counter() - t0) * 1e3)
    return {"workload": "extract_and_analyze_weights on a live ViT-Tiny-shaped module (36 matrices, result dict on the host)",
            "matrices": len(out["per_layer_metrics"]), "ms_median": statistics.median(times), "ms_min": min(times), "reps": reps}


def measure_h2d_ceiling(dev, nbytes: int, sync_all, reps: int = 10, src=None) -> float:
    """Pinned host -> device copy bandwidth (GB/s) of this rank while EVERY rank copies at the same time: the ceiling
    of the e2e arm's input path on this host (its copies have the same size; `src` = the very pinned block the e2e
    arm copies from, so allocation type and placement are the same)."""
    import torch
    import torch.distributed as dist

    if src is not None:
        h = src.view(torch.uint8)[:nbytes]
    else:
        h = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
        h.fill_(1)  # touch every page before timing
    d = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    for _ in range(3):
        d.copy_(h, non_blocking=True)
    sync_all()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        d.copy_(h, non_blocking=True)
    e1.record()
    sync_all()
    gbs = torch.tensor([nbytes * reps / 1e9 / (e0.elapsed_time(e1) / 1e3)], device=dev, dtype=torch.float64)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(gbs, op=dist.ReduceOp.MIN)  # the slowest rank bounds the step
    return float(gbs.item())


def oracle_sample_check(rec, sv, sv_offsets, host_views, n_samples: int = 64, seed: int = 0) -> dict:
    """Sampled parity check of the LAST timed step against the oracle, outside the timed region: metrics 1e-4,
    singular values 1e-5 (element-wise), integer outputs exact."""
    import numpy as np

    sys.path.insert(0, str(ROOT / "oracle"))
    import spectral_oracle as orc

    rng = np.random.default_rng(seed)
    idx = rng.choice(len(host_views), size=min(n_samples, len(host_views)), replace=False)
    worst_m = worst_sv = 0.0
    for i in idx:
        w = host_views[i]()
        ref = orc.get_spectral_metrics(w)
        for q, k in enumerate(orc.METRIC_KEYS):
            worst_m = max(worst_m, abs(float(rec["metrics"][i][q]) - ref[k]) / max(abs(ref[k]), 1e-3))
        io = orc.integer_outputs(w)
        assert (int(rec["m"][i]), int(rec["start"][i]), int(rec["end"][i]), int(rec["k"][i])) == (io["m"], io["start"], io["end"], io["k"]), i
        sref = orc.singular_values(w)
        s = sv[sv_offsets[i] : sv_offsets[i + 1]]
        worst_sv = max(worst_sv, float(np.max(np.abs(s - sref) / sref)))
    assert worst_m < 1e-4 and worst_sv < 1e-5, (worst_m, worst_sv)
    return {"samples": int(len(idx)), "max_metric_rel_err": worst_m, "max_sv_rel_err": worst_sv, "ints_exact": True,
            "gates": {"metrics": 1e-4, "sv": 1e-5}}


# ------------------------------------------------------------------------ GPU arm
def run_b200_arm(args) -> None:
    import numpy as np
    import torch
    import torch.distributed as dist

    import vision_spectra_b200 as pkg
    from vision_spectra_b200.sweep import CheckpointLayout, SweepRunner, gather_records

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # NCCL writes its version banner to stdout on first use; keep stdout for the one JSON line
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            import datetime

            dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))  # fail fast, never hang a box
            dist.barrier(device_ids=[local])
            torch.cuda.synchronize(dev)
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    lay = CheckpointLayout.vit(SCENARIO["embed_dim"], SCENARIO["depth"])
    n_ckpt = args.ckpts
    eng = pkg.SpectraEngine(dev)
    runner = SweepRunner(eng, lay, ckpts_per_chunk=args.chunk, lanes=args.lanes)

    # synthetic random-init weights, generated on the device
    arenas = []
    for c in range(n_ckpt):
        seed = SCENARIO["seeds"][c % 3] * 1_000_003 + c // 3 + 1_000 * rank
        g = torch.Generator(device=dev).manual_seed(seed)
        arenas.append(torch.randn(lay.arena_elems, generator=g, device=dev, dtype=torch.float32) * 0.02)
    matrices = n_ckpt * lay.matrices
    in_bytes = n_ckpt * lay.bytes

    def step_device():
        res = runner.run_device(arenas, want_sv=True, pipelined=args.pipelined_device)
        out = gather_records(res.records) if world > 1 else res.records
        return res, out

    def sync_all():
        if world > 1:
            dist.barrier(device_ids=[local])
        torch.cuda.synchronize(dev)

    # ---------------- device-resident throughput ("value")
    for _ in range(args.warmup):
        step_device()
    sync_all()
    eng.lib.vsp_reset_kernel_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(physical_gpu_index(local)) as clocks:
        e0.record()
        for _ in range(args.steps):
            res, gathered = step_device()
        e1.record()
        sync_all()
    launches = int(eng.lib.vsp_kernel_launch_count())
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(lt := torch.tensor([launches], device=dev), op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    ms_total = float(ms.item())
    value = world * matrices * args.steps / (ms_total / 1e3)

    # sanity: every record of the last step is a finished, finite analysis
    rec = res.records_host()
    # status 0, or 96 = ill-conditioned (kappa > 3e4) and re-solved from W; nothing else is acceptable
    assert rec.shape[0] == matrices and bool(((rec["status"] == 0) | (rec["status"] == 96)).all()), "bench records not clean"
    n_refined = int((rec["status"] == 96).sum())
    if rank == 0 and world > 1:
        assert gathered.numel() == world * matrices * 64

    # ---------------- per-stage timing (CUDA events between the kernels, same stream)
    stage_acc = np.zeros(3)
    reps = max(1, min(3, args.steps))
    for _ in range(reps):
        sm: list = []
        runner.run_device(arenas, want_sv=True, stage_ms=sm)
        stage_acc += np.array(sm)
    stage_ms = stage_acc / reps
    sync_all()

    # ---------------- end to end from pinned host memory ("e2e")
    # one pinned host block; the sweep API takes per-checkpoint arenas (views of it)
    host_block = torch.empty(n_ckpt * lay.arena_elems, dtype=torch.float32).pin_memory()
    host_arenas = []
    for c, a in enumerate(arenas):
        v = host_block[c * lay.arena_elems : (c + 1) * lay.arena_elems]
        v.copy_(a[: lay.arena_elems])
        host_arenas.append(v)
    torch.cuda.synchronize(dev)
    for _ in range(max(1, args.warmup)):
        runner.run_host(host_arenas, want_sv=True)
    sync_all()
    t0 = time.perf_counter()
    e2e_steps = max(1, args.steps)
    for _ in range(e2e_steps):
        rec_h, sv_h = runner.run_host(host_arenas, want_sv=True)
    torch.cuda.synchronize(dev)
    e2e_s = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(e2e_s, op=dist.ReduceOp.MAX)
    e2e_value = world * matrices * e2e_steps / float(e2e_s.item())
    d2h_bytes = rec_h.nbytes + (sv_h.nbytes if sv_h is not None else 0)
    assert bool(((rec_h["status"] == 0) | (rec_h["status"] == 96)).all())
    np.testing.assert_allclose(rec_h["metrics"], rec["metrics"], rtol=1e-12)  # same answers both ways
    # H2D ceiling of this host at this rank count (all ranks copy at once; same copy size as the e2e chunks)
    e2e_step_s = float(e2e_s.item()) / e2e_steps
    h2d_achieved = in_bytes / 1e9 / e2e_step_s
    h2d_ceiling = measure_h2d_ceiling(dev, min(in_bytes, args.chunk * lay.bytes), sync_all, src=host_block)

    # ---------------- sampled parity check of the last e2e step against the oracle (rank 0, outside the timed regions)
    parity = None
    if rank == 0:
        mats_per_ck = lay.matrices
        sv_off = np.zeros(matrices + 1, np.int64)
        np.cumsum(np.tile(np.array([min(sl.rows, sl.cols) for sl in lay.slots]), n_ckpt), out=sv_off[1:])

        def view(i):
            c, m_ = divmod(i, mats_per_ck)
            sl = lay.slots[m_]
            return lambda: host_arenas[c][sl.offset : sl.offset + sl.rows * sl.cols].view(sl.rows, sl.cols).numpy()

        parity = oracle_sample_check(rec_h, sv_h, sv_off, [view(i) for i in range(matrices)], n_samples=args.parity_samples)

    # ---------------- the other BASELINE.json configs and the single-checkpoint user path
    extra = {}
    if not args.no_extra:
        del host_block, host_arenas
        arenas.clear()
        torch.cuda.empty_cache()
        six = extra_six_scenario_sweep(eng, dev, world, rank, sync_all)
        if rank == 0:
            extra["six_scenario_sweep"] = six
            extra["scenario_E"] = extra_scenario(eng, dev, "E", lambda: torch.cuda.synchronize(dev))
            extra["scenario_C"] = extra_scenario(eng, dev, "C", lambda: torch.cuda.synchronize(dev))
            extra["single_checkpoint_latency"] = extra_single_checkpoint_latency(dev)
            extra["vit_base_stress"] = extra_vit_base_stress(eng, dev, lambda: torch.cuda.synchronize(dev), ckpts=args.base_ckpts)
        if world > 1:
            dist.barrier(device_ids=[local])

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---------------- roofline of the dominant kernel
    peaks = load_peaks()
    gram_name = "slice_i8_kernel+gram_i8_mma_kernel"  # stage 1: int8-split Gram on tcgen05 (26 exact digit-pair products)
    eig_name = "sbr8_kernel+chase8_kernel"  # stage 2a: blocked Householder to bandwidth 8 on DMMA.8x8x4 + bulge chasing
    names = [gram_name, eig_name, "bisect_metrics_kernel"]
    fp64_peak, fp64_src = measure_fp64_peak()
    i8_n64, i8_n256, i8_src = measure_i8_peak()
    alg = {
        # stage 1 at n = 192 with a 26-pass int8 split is tensor-bound (crossover n = 214 / passes): the
        # algorithmic 2 n^2 K flops against (measured int8 peak) / 26 digit-pair products per algorithmic one
        gram_name: {"bound": "tensor", "work": n_ckpt * lay.flops_gram() / 1e12, "unit": "TFLOP/s", "peak": i8_n256 / 26.0,
                    "flops": n_ckpt * lay.flops_gram()},
        # "tensor": the FP64 tensor cores (DMMA.8x8x4); DMMA and DFMA share the same units and the same measured peak
        eig_name: {"bound": "tensor", "work": n_ckpt * lay.flops_tridiag() / 1e12, "unit": "TFLOP/s", "peak": fp64_peak},
        "bisect_metrics_kernel": {"bound": "tensor", "work": None, "unit": "TFLOP/s", "peak": fp64_peak},
    }
    stages = []
    for nm, t in zip(names, stage_ms):
        a = alg[nm]
        ach = None if a["work"] is None else a["work"] / (t / 1e3)
        st = {"kernel": nm, "ms": float(t), "share": float(t / stage_ms.sum()), "bound": a["bound"],
              "achieved": ach, "peak": a["peak"], "unit": a["unit"],
              "frac": None if ach is None else ach / a["peak"]}
        if "flops" in a:  # the int8 split executes 26 digit-pair MMAs (x 2 K-halves of a 64-byte chunk) per algorithmic one
            st["peak_source"] = i8_src
            st["int8_peak_tops"] = {"m128n64k32 (production shape)": i8_n64, "m128n256k32": i8_n256}
            st["tensor_int8_tops_executed"] = 26 * a["flops"] / 1e12 / (t / 1e3)
            st["hbm_gbs"] = in_bytes / 1e9 / (t / 1e3)
            st["hbm_frac"] = st["hbm_gbs"] / peaks["hbm_gbs"]
        stages.append(st)
    dom = max((s for s in stages if s["achieved"] is not None), key=lambda s: s["ms"])
    roofline = {
        "kernel": dom["kernel"],
        "bound": dom["bound"],
        "achieved": dom["achieved"],
        "peak": dom["peak"],
        "unit": dom["unit"],
        "frac": dom["frac"],
        "traffic": load_traffic()[0],
        "traffic_source": load_traffic()[1],
        "peak_source": peaks["source"] if dom["bound"] == "hbm" else fp64_src,
        "algorithmic": "4*rows*cols bytes per matrix (hbm) / (4/3) n^3 flops per matrix (fp64 reduction to tridiagonal form); DESIGN.md",
        "hbm_gbs_whole_step": in_bytes / 1e9 / (ms_total / args.steps / 1e3),
        "hbm_frac_whole_step": in_bytes / 1e9 / (ms_total / args.steps / 1e3) / peaks["hbm_gbs"],
    }

    # ---------------- CPU baseline on this box (bounded sample, fresh process: no CUDA state is forked)
    cpu = None
    if not args.no_cpu_baseline:
        import subprocess

        env = {k: v for k, v in os.environ.items() if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE", "MASTER_ADDR", "MASTER_PORT")}
        env["CUDA_VISIBLE_DEVICES"] = ""
        try:
            out = subprocess.run(
                [sys.executable, str(ROOT / "bench.py"), "--impl", "reference", "--steps", "2", "--warmup", "1"],
                capture_output=True, text=True, timeout=600, env=env,
            )
            cpu = json.loads(out.stdout.strip().splitlines()[-1])["cpu_baseline"]
        except Exception as exc:  # the GPU numbers stand on their own; say why the baseline is missing
            cpu = {"value": None, "unit": "matrices/s", "cores": host_cores(), "kind": reference_kind(), "sample": f"failed: {exc!r}"}

    line = {
        "metric": "weight matrices spectrally analysed/sec (ViT-Tiny Q/K/V/MLP)",
        "value": value,
        "unit": "matrices/s",
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_total / args.steps,
        "higher_is_better": True,
        "scaling": "weak",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": SCENARIO["name"],
            "checkpoints_per_gpu": n_ckpt,
            "matrices_per_gpu_per_step": matrices,
            "input_bytes_per_gpu_per_step": in_bytes,
            "l2": "inputs (987 MB per GPU) larger than L2 (126 MB); no flush needed",
            "parallelism": f"independent checkpoints per rank x{world}; one NCCL gather of 64-byte records per step",
            "outputs": "singular values (f64) + 64-byte record per matrix",
            "refined_per_gpu_per_step": n_refined,
        },
        "e2e": {"value": e2e_value, "unit": "matrices/s", "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": int(d2h_bytes),
                "steps": e2e_steps, "api": "vision_spectra_b200.sweep.SweepRunner.run_host (pinned host arenas)",
                "h2d_gbs_achieved_per_gpu": h2d_achieved, "h2d_gbs_ceiling_per_gpu": h2d_ceiling,
                "h2d_frac": h2d_achieved / h2d_ceiling,
                "h2d_ceiling_how": "pinned host->device copies of one e2e chunk, every rank copying at the same time, slowest rank"},
        "gpu_launches": launches,
        "clocks": clocks.summary(),
        "roofline": roofline,
        "roofline_stages": stages,
        "cpu_baseline": cpu,
        "parity_check": parity,
        "extra": extra,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--ckpts", type=int, default=SCENARIO["epochs"] * len(SCENARIO["seeds"]), help="checkpoints per GPU per step")
    ap.add_argument("--chunk", type=int, default=8, help="checkpoints per H2D/compute pipeline chunk (e2e)")
    ap.add_argument("--lanes", type=int, default=6, help="compute lanes the e2e pipeline rotates chunks over")
    ap.add_argument("--pipelined-device", action="store_true",
                    help="device-resident arm: split the batch over the compute lanes too (measured slower than one launch sequence: 228k vs 243k matrices/s)")
    ap.add_argument("--ref-ckpts", type=int, default=4, help="checkpoints per step of the reference arm")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the `extra` block (other BASELINE configs, single-checkpoint latency)")
    ap.add_argument("--parity-samples", type=int, default=64, help="records of the last step checked against the oracle")
    ap.add_argument("--base-ckpts", type=int, default=16, help="ViT-Base checkpoints of the stress entry of `extra`")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        args.warmup = 3  # timing rule: W >= 3
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_b200_arm(args)


if __name__ == "__main__":
    main()
