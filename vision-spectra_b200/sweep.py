"""Checkpoint sweeps: arena layouts, host->device pipelining and multi-GPU sharding.

A sweep is "checkpoints x layers x seeds" (BASELINE.json); every matrix is an
independent work item, so ranks own disjoint shards and the only exchange is one
gather of the 64-byte result records.  The reference has no
counterpart: it analyses one live model at a time in a serial Python loop
(experiments/run_spectral_analysis.py:323, :714-717).
"""

from __future__ import annotations

from dataclasses import dataclass

import numpy as np
import torch

from . import _native as nat
from .engine import BatchResult, SpectraEngine


# --------------------------------------------------------------------- layouts
@dataclass(frozen=True)
class MatrixSlot:
    name: str  # reference naming, e.g. "blocks.0.attn.qkv.q" (extraction.py:70)
    matrix_type: str  # q | k | v | attn_proj | mlp_up | mlp_down | patch_embed
    layer_idx: int | None
    offset: int  # elements from the start of the arena
    rows: int
    cols: int


@dataclass(frozen=True)
class CheckpointLayout:
    """Where each analysed matrix of one checkpoint lives inside a flat fp32 arena.
    Slots are listed in the order `extract_and_analyze_weights` visits them
    (all q/k/v, then all attention projections, then all MLP matrices;
    run_spectral_analysis.py:313-317)."""

    slots: tuple[MatrixSlot, ...]
    arena_elems: int

    @classmethod
    def vit(cls, embed_dim: int, depth: int, mlp_ratio: int = 4, prefix: str = "") -> "CheckpointLayout":
        d, h = embed_dim, mlp_ratio * embed_dim
        per_block = 3 * d * d + d * d + h * d + d * h
        qkv, proj, mlp = [], [], []
        for b in range(depth):
            base = b * per_block
            blk = f"{prefix}blocks.{b}"
            for j, t in enumerate(("q", "k", "v")):
                qkv.append(MatrixSlot(f"{blk}.attn.qkv.{t}", t, b, base + j * d * d, d, d))
            proj.append(MatrixSlot(f"{blk}.attn.proj", "attn_proj", b, base + 3 * d * d, d, d))
            mlp.append(MatrixSlot(f"{blk}.mlp.fc1", "mlp_up", b, base + 4 * d * d, h, d))
            mlp.append(MatrixSlot(f"{blk}.mlp.fc2", "mlp_down", b, base + 4 * d * d + h * d, d, h))
        return cls(tuple(qkv + proj + mlp), depth * per_block)

    @property
    def matrices(self) -> int:
        return len(self.slots)

    @property
    def bytes(self) -> int:
        return 4 * self.arena_elems

    def views(self, arena: torch.Tensor) -> list[torch.Tensor]:
        """Zero-copy 2-D views of a flat arena tensor (q/k/v are row blocks of the
        fused qkv buffer, as in extraction.py:59-62)."""
        flat = arena.view(-1)
        return [flat[s.offset : s.offset + s.rows * s.cols].view(s.rows, s.cols) for s in self.slots]

    def flops_gram(self) -> float:
        return float(sum(2.0 * min(s.rows, s.cols) ** 2 * max(s.rows, s.cols) for s in self.slots))

    def flops_tridiag(self) -> float:
        return float(sum(4.0 / 3.0 * min(s.rows, s.cols) ** 3 for s in self.slots))


# ------------------------------------------------------------------- partitioning
def matrix_cost(rows: int, cols: int) -> float:
    """Relative cost of one work item: Gram 2 n^2 K (tensor/FP64 pipe) plus the
    eigensolve ~ 12 n^3 FP64-equivalent flops (reduction + bisection)."""
    n, k = min(rows, cols), max(rows, cols)
    return 2.0 * n * n * k + 12.0 * float(n) ** 3


def partition_lpt(costs: list[float], world_size: int, groups: list[int] | None = None) -> list[list[int]]:
    """Longest-processing-time assignment of work items to ranks.  `groups[i]` ties
    items together (q/k/v of one block share a buffer and must stay on one rank).
    Deterministic: every rank computes the same partition without communicating."""
    n = len(costs)
    if groups is None:
        groups = list(range(n))
    bundles: dict[int, list[int]] = {}
    for i, g in enumerate(groups):
        bundles.setdefault(g, []).append(i)
    order = sorted(bundles.values(), key=lambda idx: (-sum(costs[i] for i in idx), idx[0]))
    loads = [0.0] * world_size
    shards: list[list[int]] = [[] for _ in range(world_size)]
    for idx in order:
        r = min(range(world_size), key=lambda q: (loads[q], q))
        shards[r].extend(idx)
        loads[r] += sum(costs[i] for i in idx)
    for s in shards:
        s.sort()
    return shards


def shard_checkpoints(num_checkpoints: int, world_size: int, rank: int) -> list[int]:
    """Equal-cost checkpoints (same model) are dealt round-robin."""
    return list(range(rank, num_checkpoints, world_size))


# ------------------------------------------------------------------------ gather
def gather_records(local: torch.Tensor, dst: int = 0, group=None) -> torch.Tensor | None:
    """Gather fixed-size result records (uint8 tensor, 64 bytes per matrix, equal
    count on every rank) on `dst`; NCCL on CUDA tensors, gloo on CPU tensors.  Returns
    the concatenated [world*count*64] tensor on dst, None elsewhere."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if rank == dst:
        out = torch.empty(world * local.numel(), dtype=local.dtype, device=local.device)
        dist.gather(local, list(out.chunk(world)), dst=dst, group=group)
        return out
    dist.gather(local, None, dst=dst, group=group)
    return None


def gather_records_ragged(local: torch.Tensor, dst: int = 0, group=None) -> torch.Tensor | None:
    """gatherv for shards of different length: pad to the longest, gather, trim."""
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world = dist.get_world_size(group)
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    padded = torch.zeros(max(sizes), dtype=local.dtype, device=local.device)
    padded[: local.numel()] = local
    full = gather_records(padded, dst, group)
    if full is None:
        return None
    return torch.cat([c[:s] for c, s in zip(full.chunk(world), sizes)])


def records_from_bytes(buf: torch.Tensor) -> np.ndarray:
    return buf.cpu().numpy().view(nat.RECORD_DTYPE)


# ------------------------------------------------------------------------ runner
class SweepRunner:
    """Runs a list of same-layout checkpoints through one engine.

    run_device : arenas already in HBM -> one batched launch sequence.
    run_host   : arenas in pinned host memory -> chunks are copied on a side stream
                 into rotating device buffers while earlier chunks are being analysed on
                 rotating compute lanes; records and singular values come back per chunk.
    """

    def __init__(self, engine: SpectraEngine, layout: CheckpointLayout, ckpts_per_chunk: int = 8, lanes: int = 6, ramp: bool = True):
        self.engine = engine
        self.ramp = ramp
        self.layout = layout
        self.chunk = max(1, ckpts_per_chunk)
        self.nlanes = max(1, lanes)
        self._slots: list[torch.Tensor] | None = None
        self._copy_stream: torch.cuda.Stream | None = None
        # run_host alternates chunks between two compute lanes (engine + stream + workspace each),
        # so the long tail of one chunk (the FP64 re-solve of an ill-conditioned matrix keeps one SM
        # busy for milliseconds) overlaps the next chunk's kernels instead of stalling the stream.
        self._lanes: list[tuple[SpectraEngine, torch.cuda.Stream]] | None = None
        self._pinned: dict[tuple, torch.Tensor] = {}
        self._tables: dict[tuple, tuple] = {}
        self._slot_off = np.array([4 * s.offset for s in layout.slots], dtype=np.uint64)
        self._rows = np.array([s.rows for s in layout.slots], dtype=np.int32)
        self._cols = np.array([s.cols for s in layout.slots], dtype=np.int32)

    def _table(self, nck: int, want_sv: bool, engine: SpectraEngine | None = None):
        """rows / cols / ld tables and the cached plan for `nck` checkpoints."""
        engine = engine or self.engine
        key = (nck, want_sv, id(engine))
        t = self._tables.get(key)
        if t is None:
            rows, cols = np.tile(self._rows, nck), np.tile(self._cols, nck)
            ld = cols.astype(np.int64)
            plan = engine.make_plan(rows, cols, ld, nat.VSP_F32, want_sv=want_sv)
            t = self._tables[key] = (rows, cols, ld, plan)
        return t

    def _ptrs(self, bases: np.ndarray) -> np.ndarray:
        return (bases[:, None] + self._slot_off[None, :]).reshape(-1)

    def _pin(self, tag: str, numel: int, dtype) -> torch.Tensor:
        buf = self._pinned.get((tag, dtype))
        if buf is None or buf.numel() < numel:
            buf = self._pinned[(tag, dtype)] = torch.empty(numel, dtype=dtype).pin_memory()
        return buf[:numel]

    def _ensure_lanes(self):
        dev = self.engine.device
        if self._lanes is None:
            self._lanes = [(self.engine if k == 0 else SpectraEngine(dev), torch.cuda.Stream(device=dev)) for k in range(self.nlanes)]
        return self._lanes

    def run_device(self, arenas: list[torch.Tensor], want_sv: bool = True, stage_ms: list | None = None,
                   pipelined: bool = False) -> BatchResult:
        """Arenas already in HBM.  Default: all checkpoints in one launch sequence; pointer tables are built
        with NumPy (one data_ptr() per arena), not per-matrix Python.  `pipelined=True` splits the batch into
        `lanes` chunks that run on rotating compute lanes (engine + stream + workspace each), so that the tail
        of one chunk -- the FP64 re-solve of a few ill-conditioned matrices keeps a handful of SMs busy for
        milliseconds after everything else has drained -- overlaps the next chunk's (or the next call's)
        kernels; the chunks write straight into one record / singular-value buffer."""
        for a in arenas:
            if a.dtype != torch.float32 or a.device != self.engine.device or a.numel() < self.layout.arena_elems or not a.is_contiguous():
                raise ValueError("run_device: arenas must be contiguous float32 tensors of the layout's size on the engine device")
        bases = np.array([a.data_ptr() for a in arenas], dtype=np.uint64)
        nck = len(arenas)
        if not pipelined or stage_ms is not None or self.nlanes < 2 or nck < 2 * self.nlanes:
            rows, cols, ld, plan = self._table(nck, want_sv)
            return self.engine.analyze_raw(self._ptrs(bases), rows, cols, ld, nat.VSP_F32, want_sv=want_sv, plan=plan, stage_ms=stage_ms)
        dev = self.engine.device
        mats = self.layout.matrices
        n_sv = sum(min(sl.rows, sl.cols) for sl in self.layout.slots)
        records = torch.empty(nck * mats * nat.RECORD_DTYPE.itemsize, dtype=torch.uint8, device=dev)
        sv = torch.empty(nck * n_sv, dtype=torch.float64, device=dev) if want_sv else None
        main = torch.cuda.current_stream(dev)
        lanes = self._ensure_lanes()
        per = (nck + self.nlanes - 1) // self.nlanes
        for li, c0 in enumerate(range(0, nck, per)):
            lane_eng, lane_stream = lanes[li % self.nlanes]
            cnt = min(per, nck - c0)
            lane_stream.wait_stream(main)
            with torch.cuda.stream(lane_stream):
                rows, cols, ld, plan = self._table(cnt, want_sv, lane_eng)
                lane_eng.analyze_raw(self._ptrs(bases[c0 : c0 + cnt]), rows, cols, ld, nat.VSP_F32, want_sv=want_sv, plan=plan,
                                     out_records=records[c0 * mats * 64 : (c0 + cnt) * mats * 64],
                                     out_sv=None if sv is None else sv[c0 * n_sv : (c0 + cnt) * n_sv])
                if c0:  # records carry chunk-local item ids: first int32 of every 64-byte record
                    records[c0 * mats * 64 : (c0 + cnt) * mats * 64].view(torch.int32).view(-1, 16)[:, 0] += c0 * mats
        for _, cs in lanes:
            main.wait_stream(cs)
        rows_all, cols_all = np.tile(self._rows, nck), np.tile(self._cols, nck)
        offs = np.zeros(nck * mats + 1, np.int64)
        np.cumsum(np.minimum(rows_all, cols_all), out=offs[1:])
        return BatchResult(records, sv, offs, nck * mats)

    def _chunk_starts(self, nck: int) -> list[int]:
        """Chunk boundaries of run_host: ramp up 1/4, 1/2 of a chunk, full chunks, ramp down 1/2, 1/4, 1/8."""
        c = self.chunk
        if not self.ramp or nck < 4 * c or c < 4:
            return list(range(0, nck, c)) + [nck]
        head, tail = [c // 4, c // 2], [c // 2, c // 4]
        if c >= 8:  # the step ends one chunk latency (~2 ms) after the last byte has landed: make that chunk a small one
            tail.append(c // 8)
        mid = nck - sum(head) - sum(tail)
        full, rem = divmod(mid, c)
        body = [c] * full
        if rem and full:  # no tiny remainder chunk: split (chunk + remainder) evenly instead
            body[-1:] = [(c + rem + 1) // 2, (c + rem) // 2]
        elif rem:
            body = [rem]
        sizes = head + body + tail
        starts = [0]
        for sz in sizes:
            starts.append(starts[-1] + sz)
        return starts

    def run_host(self, arenas: list[torch.Tensor], want_sv: bool = True) -> tuple[np.ndarray, np.ndarray | None]:
        eng, lay = self.engine, self.layout
        dev = eng.device
        nck = len(arenas)
        mats = lay.matrices
        n_sv = sum(min(s.rows, s.cols) for s in lay.slots)
        rec_host = self._pin("rec", nck * mats * 64, torch.uint8)
        sv_host = self._pin("sv", nck * n_sv, torch.float64) if want_sv else None
        if self._slots is None or self._slots[0].numel() < self.chunk * lay.arena_elems:
            self._slots = [torch.empty(self.chunk * lay.arena_elems, dtype=torch.float32, device=dev) for _ in range(self.nlanes)]
            self._copy_stream = torch.cuda.Stream(device=dev)
        main = torch.cuda.current_stream(dev)
        copy = self._copy_stream
        self._ensure_lanes()
        copy.wait_stream(main)
        for _, cs in self._lanes:
            cs.wait_stream(main)
        L = self.nlanes
        free = [None] * L  # event: slot's previous consumer finished
        keep = []
        # results leave the pinned staging buffers chunk by chunk while later chunks are still in flight: only the last
        # (short) chunks are copied out after the device has drained
        rec = np.empty(nck * mats, dtype=nat.RECORD_DTYPE)
        sv_out = np.empty(nck * n_sv, dtype=np.float64) if want_sv else None
        rec_view = rec_host.numpy().view(nat.RECORD_DTYPE)
        sv_view = sv_host.numpy() if want_sv else None
        landed: list[tuple] = []  # (event after the chunk's D2H copies, c0, c1), in launch order

        def unload(c0: int, c1: int) -> None:
            rec[c0 * mats : c1 * mats] = rec_view[c0 * mats : c1 * mats]
            rec["item"][c0 * mats : c1 * mats] += c0 * mats  # records carry chunk-local item ids
            if want_sv:
                sv_out[c0 * n_sv : c1 * n_sv] = sv_view[c0 * n_sv : c1 * n_sv]

        # chunk schedule: short chunks at both ends (the first chunk's copy and the last chunk's kernels are the only
        # parts of the step that do not overlap anything), full ones in between
        starts = self._chunk_starts(nck)
        for ci, (c0, c1) in enumerate(zip(starts[:-1], starts[1:])):
            chunk = arenas[c0:c1]
            slot = self._slots[ci % L]
            lane_eng, lane_stream = self._lanes[ci % L]
            with torch.cuda.stream(copy):
                if free[ci % L] is not None:
                    copy.wait_event(free[ci % L])
                # consecutive views of one pinned host block go over the bus as ONE copy (55 GB/s measured for
                # large copies vs 48 GB/s for 10.6 MB ones); separate tensors as one copy each
                j = 0
                while j < len(chunk):
                    a = chunk[j]
                    k = 1
                    base = a._base
                    if base is not None and base.dim() == 1 and a.is_contiguous():
                        while (j + k < len(chunk) and chunk[j + k]._base is base and chunk[j + k].is_contiguous()
                               and chunk[j + k].storage_offset() == a.storage_offset() + k * lay.arena_elems
                               and a.numel() == lay.arena_elems):
                            k += 1
                    if k > 1:
                        o = a.storage_offset() - base.storage_offset()
                        slot[j * lay.arena_elems : (j + k) * lay.arena_elems].copy_(base[o : o + k * lay.arena_elems], non_blocking=True)
                    else:
                        slot[j * lay.arena_elems : (j + 1) * lay.arena_elems].copy_(a.view(-1), non_blocking=True)
                    j += k
                ready = torch.cuda.Event()
                ready.record(copy)
            with torch.cuda.stream(lane_stream):
                lane_stream.wait_event(ready)
                rows, cols, ld, plan = self._table(len(chunk), want_sv, lane_eng)
                bases = np.uint64(slot.data_ptr()) + np.arange(len(chunk), dtype=np.uint64) * np.uint64(lay.bytes)
                res = lane_eng.analyze_raw(self._ptrs(bases), rows, cols, ld, nat.VSP_F32, want_sv=want_sv, plan=plan)
                done = torch.cuda.Event()
                done.record(lane_stream)
                free[ci % L] = done
                r0 = c0 * mats * 64
                rec_host[r0 : r0 + res.records.numel()].copy_(res.records, non_blocking=True)
                if want_sv:
                    s0 = c0 * n_sv
                    sv_host[s0 : s0 + res.sv.numel()].copy_(res.sv, non_blocking=True)
                home = torch.cuda.Event()
                home.record(lane_stream)
                landed.append((home, c0, c1))
            keep.append(res)
            while landed and landed[0][0].query():
                _, a0, a1 = landed.pop(0)
                unload(a0, a1)
        for home, a0, a1 in landed:
            home.synchronize()
            unload(a0, a1)
        for _, cs in self._lanes:
            main.wait_stream(cs)
        return rec, sv_out
