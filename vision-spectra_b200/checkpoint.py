"""Checkpoint -> device-batch feeder.

The reference analyses a *live* timm model and copies every module's weight to the host
one by one (metrics/extraction.py:56,101,142,184,226).  For sweeps over saved checkpoints
(`torch.save({"epoch", "model_state_dict", ...})`, training/base.py:576-594;
utils/checkpointing.py:15-57) this module selects the same matrices straight from the
state dict -- same names, order, matrix types and layer indices as extract_qkv_weights /
extract_attention_weights / extract_mlp_weights / extract_patch_embed_weights -- without
instantiating timm, packs them into one pinned arena, uploads them with a single copy and
returns device views (q/k/v are row blocks of the fused qkv buffer).
"""

from __future__ import annotations

from pathlib import Path
from typing import Any, Mapping

import numpy as np
import torch

from .engine import analyze_matrices
from .metrics.extraction import WeightInfo, _extract_layer_idx, _selected
from .metrics.spectral import aggregate_spectral_metrics


def unwrap_state_dict(obj: Any) -> Mapping[str, torch.Tensor]:
    """Accept a raw state dict or the reference's checkpoint dict (training/base.py:576-594)."""
    if isinstance(obj, Mapping):
        for key in ("model_state_dict", "state_dict", "model"):
            if key in obj and isinstance(obj[key], Mapping):
                return obj[key]
        return obj
    raise TypeError(f"not a state dict: {type(obj)}")


def _module_names(sd: Mapping[str, torch.Tensor]) -> list[str]:
    """Every module path implied by the parameter names, parents before children, in
    state-dict (= module registration) order: the order `named_modules()` would give."""
    seen: dict[str, None] = {"": None}
    for key in sd:
        parts = key.split(".")[:-1]
        for i in range(1, len(parts) + 1):
            seen.setdefault(".".join(parts[:i]), None)
    return list(seen)


def _has(sd, name: str) -> bool:
    t = sd.get(name)
    return isinstance(t, torch.Tensor)


def _join(mod: str, leaf: str) -> str:
    return f"{mod}.{leaf}" if mod else leaf


def select_matrices(
    state_dict: Mapping[str, torch.Tensor],
    layer_patterns: list[str] | None = None,
    include_qkv: bool = True,
    include_proj: bool = True,
    include_mlp: bool = True,
    include_patch_embed: bool = False,
) -> list[WeightInfo]:
    """Same selection as the reference's extractors, applied to parameter names.  Defaults
    reproduce `extract_and_analyze_weights` (qkv + attention proj + mlp,
    experiments/run_spectral_analysis.py:313-317); pass the `SpectralConfig` flags for the
    trainer's union (metrics/extraction.py:245-281)."""
    sd = unwrap_state_dict(state_dict)
    mods = _module_names(sd)
    out: list[WeightInfo] = []

    def info(name, layer_idx, mtype, w):
        return WeightInfo(name=name, layer_idx=layer_idx, matrix_type=mtype, weight=w, shape=tuple(w.shape))

    if include_qkv:  # extraction.py:49-110
        for mod in mods:
            if not _selected(mod, layer_patterns):
                continue
            if _has(sd, _join(mod, "qkv.weight")):
                qkv = sd[_join(mod, "qkv.weight")]
                d = qkv.shape[1]
                li = _extract_layer_idx(mod)
                out.append(info(f"{mod}.qkv.q", li, "q", qkv[:d]))
                out.append(info(f"{mod}.qkv.k", li, "k", qkv[d : 2 * d]))
                out.append(info(f"{mod}.qkv.v", li, "v", qkv[2 * d :]))
            elif _has(sd, _join(mod, "q_proj.weight")):
                li = _extract_layer_idx(mod)
                for pn, pt in (("q_proj", "q"), ("k_proj", "k"), ("v_proj", "v")):
                    if _has(sd, _join(mod, f"{pn}.weight")):
                        out.append(info(f"{mod}.{pn}", li, pt, sd[_join(mod, f"{pn}.weight")]))
    if include_proj:  # extraction.py:131-153
        for mod in mods:
            if not _selected(mod, layer_patterns):
                continue
            low = mod.lower()
            if _has(sd, _join(mod, "proj.weight")) and ("attn" in low or "attention" in low):
                out.append(info(f"{mod}.proj", _extract_layer_idx(mod), "attn_proj", sd[_join(mod, "proj.weight")]))
    if include_mlp:  # extraction.py:174-203
        for mod in mods:
            if not mod or not _selected(mod, layer_patterns):
                continue
            low = mod.lower()
            if ("mlp" in low or "ffn" in low) and _has(sd, _join(mod, "weight")):
                last = mod.split(".")[-1]
                if "fc1" in mod or "0" in last:
                    mtype = "mlp_up"
                elif "fc2" in mod or "2" in last:
                    mtype = "mlp_down"
                else:
                    mtype = "mlp"
                out.append(info(mod, _extract_layer_idx(mod), mtype, sd[_join(mod, "weight")]))
    if include_patch_embed:  # extraction.py:220-240 (ignores layer_patterns)
        for mod in mods:
            if "patch_embed" in mod.lower() and _has(sd, _join(mod, "proj.weight")):
                w = sd[_join(mod, "proj.weight")]
                if w.ndim == 4:
                    w = w.reshape(w.shape[0], -1)
                out.append(info(f"{mod}.proj", None, "patch_embed", w))
    return out


def upload_matrices(infos: list[WeightInfo], device: torch.device) -> list[WeightInfo]:
    """One pinned arena, one H2D copy; fused qkv buffers are uploaded once and q/k/v stay
    views of it.  fp32 on the device (bf16/fp16 checkpoints widen exactly)."""
    bases: dict[tuple[int, int], tuple[torch.Tensor, int]] = {}  # storage ptr -> (base tensor, arena offset)
    total = 0

    def base_of(w: torch.Tensor) -> torch.Tensor:
        # a view is served from its base only if the base is laid out densely (then every strided view of it --
        # row blocks, column blocks, transposes -- can be rebuilt on the device copy); anything else is
        # uploaded as its own dense matrix
        b = w._base if w._base is not None else w
        return b if b.is_contiguous() else w.contiguous()

    resolved = [base_of(wi.weight) for wi in infos]
    for base in resolved:
        key = (base.untyped_storage().data_ptr(), base.storage_offset())
        if key not in bases:
            bases[key] = (base, total)
            total += base.numel()
    arena = torch.empty(max(total, 1), dtype=torch.float32).pin_memory()
    for base, off in bases.values():
        arena[off : off + base.numel()].copy_(base.detach().reshape(-1).to(torch.float32))
    dev = arena.to(device, non_blocking=True)
    out = []
    for wi, base in zip(infos, resolved):
        _, off = bases[(base.untyped_storage().data_ptr(), base.storage_offset())]
        w = wi.weight
        if w._base is not None and base is w._base:
            # same sizes / strides / offset as on the host, relative to the (dense) base: exact for any view
            view = torch.as_strided(dev, tuple(w.shape), tuple(w.stride()), off + w.storage_offset() - base.storage_offset())
        else:
            view = dev[off : off + base.numel()].view(base.shape)
        out.append(WeightInfo(wi.name, wi.layer_idx, wi.matrix_type, view, tuple(view.shape)))
    return out


def load_checkpoint(path: str | Path) -> Mapping[str, torch.Tensor]:
    return unwrap_state_dict(torch.load(str(path), map_location="cpu", weights_only=True))


def analyze_state_dict(state_dict: Mapping[str, torch.Tensor], device: torch.device | None = None, **select_kw) -> dict[str, Any]:
    """`extract_and_analyze_weights` (run_spectral_analysis.py:297-345) for a saved checkpoint:
    same three-key result dict, no model object needed."""
    device = device or torch.device("cuda", torch.cuda.current_device())
    infos = upload_matrices(select_matrices(state_dict, **select_kw), device)
    metrics, svs = analyze_matrices([w.weight for w in infos]) if infos else ([], [])
    per_layer, sv_out, mlist = {}, {}, []
    for w, m, s in zip(infos, metrics, svs):
        per_layer[w.name] = m
        mlist.append(m)
        sv_out[w.name] = [] if s is None else s.tolist()
    return {"per_layer_metrics": per_layer, "aggregated_metrics": aggregate_spectral_metrics(mlist), "singular_values": sv_out}


def analyze_checkpoint(path: str | Path, device: torch.device | None = None, **select_kw) -> dict[str, Any]:
    return analyze_state_dict(load_checkpoint(path), device, **select_kw)
