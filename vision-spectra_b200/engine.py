"""Batched entry point of the B200 spectral path.

`SpectraEngine.analyze(list_of_matrices)` is the one call that replaces the
reference's per-matrix Python loop (experiments/run_spectral_analysis.py:323-336,
training/base.py:399-405): every matrix of a checkpoint goes to the GPU in one
ragged batch and comes back as one 64-byte record (+ singular values).

PyTorch is used for device memory and streams only; all arithmetic happens in
lib/libvspectra.so (hand-written sm_100a CUDA behind the C-ABI of
include/vspectra.h).  There is no CPU fallback: without the library or without a
CUDA device this module raises.
"""

from __future__ import annotations

import ctypes
import os
from collections import OrderedDict
from typing import Any, Sequence

import numpy as np
import torch

from . import _native as nat

METRIC_KEYS = ("spectral_entropy", "stable_rank", "alpha_exponent", "pl_alpha_hill")
_NAN = float("nan")


def nan_metrics() -> dict[str, float]:
    """The reference's failure value: every metric NaN (spectral.py:87-93)."""
    return dict.fromkeys(METRIC_KEYS, _NAN)


class BatchResult:
    """Records + singular values of one executed batch, still on the device."""

    __slots__ = ("records", "sv", "sv_offsets", "count", "dist")

    def __init__(self, records: torch.Tensor, sv: torch.Tensor | None, sv_offsets: np.ndarray, count: int,
                 dist: torch.Tensor | None = None):
        self.records = records  # uint8 [count*64]
        self.sv = sv  # float64 [sum n] or None
        self.sv_offsets = sv_offsets  # int64 [count+1]
        self.count = count
        self.dist = dist  # float64 [count, VSP_AUX_STRIDE] or None: 4 x dist_k distribution rows, then the Clauset block

    def dist_host(self) -> np.ndarray | None:
        return None if self.dist is None else self.dist.cpu().numpy()

    def records_host(self) -> np.ndarray:
        return self.records.cpu().numpy().view(nat.RECORD_DTYPE)

    def sv_host(self) -> np.ndarray | None:
        return None if self.sv is None else self.sv.cpu().numpy()


class Plan:
    """Owner of one native `vsp_plan` (validated, device-resident shape table of a batch).

    The engine's LRU cache and every caller that keeps a plan (`SweepRunner`) hold references to this object;
    the native plan is destroyed when the last reference goes away, after the device has finished with its item
    table.  A plan must not be executed on two streams at the same time (the native side rewrites the item
    pointers and shares one fork/join event pair per plan): `SweepRunner` keeps one plan per compute lane."""

    __slots__ = ("handle", "_lib", "_device", "__weakref__")

    def __init__(self, lib, handle: int, device: torch.device):
        self._lib, self.handle, self._device = lib, handle, device

    def close(self) -> None:
        h, self.handle = self.handle, None
        if h is not None:
            try:
                torch.cuda.synchronize(self._device)  # kernels may still read the item table
            finally:
                self._lib.vsp_plan_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SpectraEngine:
    """Owns the per-device scratch (workspace, result buffers, cached plans)."""

    def __init__(self, device: torch.device | str | int | None = None, max_cached_plans: int = 8):
        if not torch.cuda.is_available():
            raise nat.NativeError("no CUDA device: the spectral path has no CPU fallback")
        self.lib = nat.load()
        if device is None:
            device = torch.device("cuda", torch.cuda.current_device())
        self.device = torch.device(device) if not isinstance(device, int) else torch.device("cuda", device)
        if self.device.type != "cuda":
            raise nat.NativeError(f"SpectraEngine needs a CUDA device, got {self.device}")
        if self.device.index is None:
            self.device = torch.device("cuda", torch.cuda.current_device())
        # analyze_device goes through the registered torch op when lib/vspectra_torch.so is there (it always is after
        # build()); VSP_NO_TORCH_OP=1 keeps the ctypes route (same kernels, same results: tests compare the two)
        self._use_torch_op = nat.TORCH_EXT_PATH.exists() and not os.environ.get("VSP_NO_TORCH_OP")
        self._plans: OrderedDict[tuple, Plan] = OrderedDict()
        self._max_plans = max_cached_plans
        self._ws: torch.Tensor | None = None
        self._pinned: torch.Tensor | None = None
        self._pinned_busy: torch.cuda.Event | None = None

    # ------------------------------------------------------------------ plans
    def _plan(self, rows, cols, ld, dtype: int, opts: nat.VspOpts) -> Plan:
        key = (rows.tobytes(), cols.tobytes(), ld.tobytes(), dtype, opts.fit_start, opts.fit_end, opts.hill_k, opts.want_sv, opts.refine, opts.dist_k, opts.clauset)
        plan = self._plans.get(key)
        if plan is not None:
            self._plans.move_to_end(key)
            return plan
        handle = ctypes.c_void_p()
        with torch.cuda.device(self.device):
            nat.check(
                self.lib.vsp_plan_create(len(rows), nat.p32(rows), nat.p32(cols), nat.p64(ld), dtype, ctypes.byref(opts), ctypes.byref(handle)),
                "vsp_plan_create",
            )
        plan = self._plans[key] = Plan(self.lib, handle.value, self.device)
        while len(self._plans) > self._max_plans:
            self._plans.popitem(last=False)  # destroyed when its last holder (e.g. a SweepRunner) lets go
        return plan

    def close(self) -> None:
        """Drop the cached plans (each is destroyed once nobody else holds it)."""
        self._plans.clear()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --------------------------------------------------------------- device path
    def analyze_device(
        self,
        tensors: Sequence[torch.Tensor],
        fit_range: tuple[int, int] | None = None,
        hill_k: int | None = None,
        want_sv: bool = True,
        dist_k: int = 0,
        clauset: bool = False,
    ) -> BatchResult:
        """Launch the three stages for 2-D CUDA tensors of ONE dtype (float32 or
        float64) with unit column stride.  Asynchronous on the current stream;
        the returned buffers are valid once that stream reaches this point."""
        count = len(tensors)
        if count == 0:
            return BatchResult(torch.empty(0, dtype=torch.uint8, device=self.device), None, np.zeros(1, np.int64), 0)
        dt = tensors[0].dtype
        if dt not in (torch.float32, torch.float64):
            raise TypeError(f"analyze_device takes float32/float64 tensors, got {dt}")
        if self._use_torch_op:
            # the PyTorch extension layer: torch.ops.vision_spectra_b200.analyze_batch (csrc/torch_ext.cpp) validates
            # the tensors, allocates the outputs with ATen and launches on the current stream under a device guard
            for t in tensors:
                if t.device != self.device:
                    raise ValueError("analyze_device: tensors must be on the engine device")
            fs, fe = (-1, -1) if fit_range is None else (int(fit_range[0]), int(fit_range[1]))
            records, sv, dist = nat.load_torch_ext().analyze_batch(list(tensors), fs, fe, -1 if hill_k is None else int(hill_k),
                                                                   bool(want_sv), int(dist_k or 0), bool(clauset))
            offs = np.zeros(count + 1, np.int64)
            np.cumsum([min(t.shape) for t in tensors], out=offs[1:])
            return BatchResult(records.view(-1), sv if want_sv else None, offs, count, dist if (dist_k or clauset) else None)
        rows = np.empty(count, np.int32)
        cols = np.empty(count, np.int32)
        ld = np.empty(count, np.int64)
        ptrs = np.empty(count, np.uint64)
        for i, t in enumerate(tensors):
            if t.dtype != dt or t.ndim != 2 or t.device != self.device or (t.shape[1] > 1 and t.stride(1) != 1):
                raise ValueError("analyze_device: tensors must be 2-D, same dtype, on the engine device, unit column stride")
            rows[i], cols[i] = t.shape
            ld[i] = t.stride(0) if t.shape[0] > 1 else max(t.shape[1], t.stride(0))
            if ld[i] < cols[i]:
                raise ValueError("analyze_device: overlapping rows (stride(0) < cols)")
            ptrs[i] = t.data_ptr()
        dtype = nat.VSP_F32 if dt == torch.float32 else nat.VSP_F64
        return self.analyze_raw(ptrs, rows, cols, ld, dtype, fit_range, hill_k, want_sv, dist_k=dist_k, clauset=clauset)

    def analyze_raw(
        self,
        ptrs: np.ndarray,
        rows: np.ndarray,
        cols: np.ndarray,
        ld: np.ndarray,
        dtype: int,
        fit_range: tuple[int, int] | None = None,
        hill_k: int | None = None,
        want_sv: bool = True,
        plan: Plan | None = None,
        stage_ms: list | None = None,
        out_records: torch.Tensor | None = None,
        out_sv: torch.Tensor | None = None,
        dist_k: int = 0,
        clauset: bool = False,
    ) -> BatchResult:
        """Table form of analyze_device: `ptrs` is a uint64 array of device addresses,
        rows/cols int32, ld int64 (elements).  The caller keeps the memory alive until
        the stream has passed this point.  Passing a `plan` (from `make_plan`) skips the
        shape-table lookup for repeated batches."""
        count = int(len(ptrs))
        ptrs = np.ascontiguousarray(ptrs, dtype=np.uint64)
        if plan is None:
            plan = self.make_plan(rows, cols, ld, dtype, fit_range, hill_k, want_sv, dist_k, clauset)
        if plan.handle is None:
            raise nat.NativeError("analyze_raw: the plan has been closed")
        ws_bytes = self.lib.vsp_plan_workspace_bytes(plan.handle)
        sv_count = self.lib.vsp_plan_sv_count(plan.handle)
        if self._ws is None or self._ws.numel() < ws_bytes:
            self._ws = None
            self._ws = torch.empty(int(ws_bytes), dtype=torch.uint8, device=self.device)
        # caller-provided output slices (a pipelined sweep writes its chunks into one buffer) or fresh ones
        records = out_records if out_records is not None else torch.empty(count * nat.RECORD_DTYPE.itemsize, dtype=torch.uint8, device=self.device)
        sv = (out_sv if out_sv is not None else torch.empty(int(sv_count), dtype=torch.float64, device=self.device)) if want_sv else None
        if records.numel() != count * nat.RECORD_DTYPE.itemsize or (sv is not None and sv.numel() != int(sv_count)):
            raise ValueError("analyze_raw: output buffers do not match the batch")
        stream = torch.cuda.current_stream(self.device).cuda_stream
        args = (
            plan.handle,
            ptrs.ctypes.data_as(ctypes.POINTER(ctypes.c_void_p)),
            None if sv is None else sv.data_ptr(),
            records.data_ptr(),
            self._ws.data_ptr(),
            int(self._ws.numel()),
            stream,
        )
        dist = None
        if dist_k or clauset:
            dist = torch.empty((count, nat.aux_stride(dist_k, clauset)), dtype=torch.float64, device=self.device)
            with torch.cuda.device(self.device):
                nat.check(self.lib.vsp_plan_execute_dist(*args[:4], dist.data_ptr(), *args[4:]), "vsp_plan_execute_dist")
            offs = np.zeros(count + 1, np.int64)
            np.cumsum(np.minimum(rows, cols), out=offs[1:])
            return BatchResult(records, sv, offs, count, dist)
        with torch.cuda.device(self.device):
            if stage_ms is None:
                nat.check(self.lib.vsp_plan_execute(*args), "vsp_plan_execute")
            else:  # per-stage CUDA-event timing (synchronises); used by bench.py's roofline
                ms = (ctypes.c_float * 3)()
                nat.check(self.lib.vsp_plan_execute_profiled(*args, ms), "vsp_plan_execute_profiled")
                stage_ms[:] = [float(ms[0]), float(ms[1]), float(ms[2])]
        offs = np.zeros(count + 1, np.int64)
        np.cumsum(np.minimum(rows, cols), out=offs[1:])
        return BatchResult(records, sv, offs, count)

    def make_plan(self, rows, cols, ld, dtype: int, fit_range=None, hill_k=None, want_sv: bool = True, dist_k: int = 0,
                  clauset: bool = False) -> Plan:
        """Validated, device-resident shape table for a batch (cached per engine; the returned `Plan` stays
        valid for as long as the caller holds it, whatever the cache evicts)."""
        opts = nat.VspOpts.make(fit_range, hill_k, want_sv, dist_k=dist_k, clauset=clauset)
        return self._plan(nat.i32(rows), nat.i32(cols), nat.i64(ld), dtype, opts)

    # ----------------------------------------------------------------- host path
    def _stage_host(self, arrays: list[np.ndarray], np_dtype) -> list[torch.Tensor]:
        """Copy host matrices through one pinned arena and one H2D transfer."""
        sizes = [a.size for a in arrays]
        total = int(sum(sizes))
        tdt = torch.float32 if np_dtype == np.float32 else torch.float64
        nbytes = total * np.dtype(np_dtype).itemsize
        if self._pinned_busy is not None:  # the previous H2D copy may still be reading the arena
            self._pinned_busy.synchronize()
        if self._pinned is None or self._pinned.numel() < nbytes:
            self._pinned = torch.empty(max(nbytes, 1 << 20), dtype=torch.uint8).pin_memory()
        arena = self._pinned[:nbytes].view(tdt)
        arena_np = arena.numpy()
        off = 0
        for a, sz in zip(arrays, sizes):
            arena_np[off : off + sz] = np.asarray(a, dtype=np_dtype).reshape(-1)
            off += sz
        dev = arena.to(self.device, non_blocking=True)
        self._pinned_busy = torch.cuda.Event()
        self._pinned_busy.record(torch.cuda.current_stream(self.device))
        out, off = [], 0
        for a, sz in zip(arrays, sizes):
            out.append(dev[off : off + sz].view(a.shape))
            off += sz
        return out

    def analyze(
        self,
        matrices: Sequence[Any],
        fit_range: tuple[int, int] | None = None,
        hill_k: int | None = None,
        want_sv: bool = True,
        dist_k: int = 0,
        dist_out: list | None = None,
        clauset_out: list | None = None,
    ) -> tuple[list[dict[str, float]], list[np.ndarray | None], np.ndarray]:
        """Analyse a ragged list of matrices (torch tensors on any device or NumPy
        arrays, any float dtype).  Returns (metrics dicts, singular-value arrays,
        records) in input order.  Anything the reference would answer with NaN
        (non-2-D input, empty matrix, NaN/Inf entries) yields NaN metrics and
        `None` singular values; nothing raises for bad *data*.

        `dist_k > 0` asks for the device-computed distribution arrays (spectral.py:545-557) truncated to `dist_k`
        entries (SpectralTracker's max_singular_values); they are appended to the list `dist_out`, one [4, min(n, dist_k)]
        array per matrix (None where the reference returns no distribution).  Passing a list as `clauset_out` switches
        on the Clauset-Shalizi-Newman x_min scan (an extra: the reference has none) and fills it with one dict per
        matrix: alpha, xmin, ks_distance, xmin_index (0 = largest eigenvalue), tail_count.

        float32 / float16 / bfloat16 inputs are analysed from their (exactly widened) float32 values; float64,
        integer and extended-precision inputs from float64, as the reference's `np.asarray(w, dtype=np.float64)`
        sees them (spectral.py:407)."""
        count = len(matrices)
        metrics: list[dict[str, float]] = [nan_metrics() for _ in range(count)]
        svs: list[np.ndarray | None] = [None] * count
        records = np.zeros(count, nat.RECORD_DTYPE)
        records["item"] = np.arange(count)
        records["status"] = nat.ST_NONFINITE
        records["start"] = records["end"] = records["k"] = -1
        records["metrics"] = np.nan
        groups: dict[Any, list[tuple[int, Any]]] = {torch.float32: [], torch.float64: []}
        host: dict[Any, list[tuple[int, np.ndarray]]] = {np.float32: [], np.float64: []}
        for i, w in enumerate(matrices):
            if isinstance(w, torch.Tensor):
                w = w.detach()
                if w.ndim != 2 or w.numel() == 0:
                    continue
                if w.device.type == "cuda":
                    if w.device != self.device:
                        w = w.to(self.device)
                    if w.dtype not in (torch.float32, torch.float64):
                        w = w.float() if w.dtype in (torch.float16, torch.bfloat16) else w.double()
                    if (w.shape[1] > 1 and w.stride(1) != 1) or (w.shape[0] > 1 and w.stride(0) < w.shape[1]):
                        w = w.contiguous()
                    groups[w.dtype].append((i, w))
                    continue
                if w.dtype not in (torch.float32, torch.float64):
                    w = w.float() if w.dtype in (torch.float16, torch.bfloat16) else w.double()
                w = w.numpy()
            else:
                w = np.asarray(w)
                if w.dtype.kind not in "fiub":
                    continue
            if w.ndim != 2 or w.size == 0:
                continue
            # half / single precision widen exactly to fp32; ints and long doubles go to float64 like the reference
            npdt = np.float32 if (w.dtype.kind == "f" and w.dtype.itemsize <= 4) else np.float64
            host[npdt].append((i, w))
        for npdt, lst in host.items():
            if lst:
                staged = self._stage_host([w for _, w in lst], npdt)
                tdt = torch.float32 if npdt == np.float32 else torch.float64
                groups[tdt] += [(i, t) for (i, _), t in zip(lst, staged)]
        pending = []
        for tdt, lst in groups.items():
            if lst:
                lst.sort(key=lambda p: p[0])
                res = self.analyze_device([t for _, t in lst], fit_range, hill_k, want_sv, dist_k, clauset_out is not None)
                pending.append((lst, res))
        dists: list = [None] * count
        clausets: list = [None] * count
        for lst, res in pending:
            rec = res.records_host()
            sv = res.sv_host()
            dh = res.dist_host()
            dk = max(0, int(dist_k or 0))
            for j, (i, t) in enumerate(lst):
                if dh is not None and dk and not (int(rec[j]["status"]) & nat.ST_NONFINITE):
                    dists[i] = dh[j][: 4 * dk].reshape(4, dk)[:, : min(dk, min(t.shape))].copy()
                if dh is not None and clauset_out is not None:
                    c = dh[j][4 * dk : 4 * dk + 8]
                    clausets[i] = {"alpha": float(c[0]), "xmin": float(c[1]), "ks_distance": float(c[2]),
                                   "xmin_index": int(c[3]), "tail_count": int(c[4])}
                r = rec[j]
                records[i] = r
                records[i]["item"] = i
                metrics[i] = {k: float(r["metrics"][q]) for q, k in enumerate(METRIC_KEYS)}
                if sv is not None and not (int(r["status"]) & nat.ST_NONFINITE):
                    svs[i] = sv[res.sv_offsets[j] : res.sv_offsets[j + 1]].copy()
        if dist_out is not None:
            dist_out[:] = dists
        if clauset_out is not None:
            clauset_out[:] = clausets
        return metrics, svs, records


_default_engines: dict[int, SpectraEngine] = {}


def default_engine(device: torch.device | None = None) -> SpectraEngine:
    """One engine per CUDA device of this process."""
    if not torch.cuda.is_available():
        raise nat.NativeError("no CUDA device: the spectral path has no CPU fallback")
    idx = torch.cuda.current_device() if device is None or device.index is None else device.index
    eng = _default_engines.get(idx)
    if eng is None:
        eng = _default_engines[idx] = SpectraEngine(torch.device("cuda", idx))
    return eng


def analyze_matrices(matrices: Sequence[Any], **kw) -> tuple[list[dict[str, float]], list[np.ndarray | None]]:
    """SURVEY 8b's batched entry: list of matrices -> (list of metric dicts, list of SV arrays)."""
    device = None
    for w in matrices:
        if isinstance(w, torch.Tensor) and w.device.type == "cuda":
            device = w.device
            break
    metrics, svs, _ = default_engine(device).analyze(matrices, **kw)
    return metrics, svs
