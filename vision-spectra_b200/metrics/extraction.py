"""Drop-in for `vision_spectra.metrics.extraction`: same selection rules, names,
order and `WeightInfo` fields, but weights stay where they are.

The reference copies every module's weight to the host (`.detach().cpu().numpy()`,
extraction.py:56,101,142,184,226).  Here `WeightInfo.weight` is a `torch.Tensor`
*view* of the live parameter (q/k/v are row-block views of the fused `[3d, d]`
qkv buffer, extraction.py:59-62), so a CUDA model is analysed with zero copies and
a CPU model is uploaded once per batch by the engine.  `WeightInfo.numpy()` gives
the reference's ndarray on demand.
"""

from __future__ import annotations

import re
from dataclasses import dataclass

import numpy as np
import torch
import torch.nn as nn

__all__ = [
    "WeightInfo",
    "extract_qkv_weights",
    "extract_attention_weights",
    "extract_mlp_weights",
    "extract_patch_embed_weights",
    "extract_all_weights",
    "group_weights_by_layer",
    "group_weights_by_type",
]


@dataclass
class WeightInfo:
    """Reference extraction.py:18-29; `weight` is a tensor view instead of an ndarray."""

    name: str
    layer_idx: int | None
    matrix_type: str
    weight: torch.Tensor
    shape: tuple[int, ...]

    def numpy(self) -> np.ndarray:
        return self.weight.detach().cpu().numpy()

    def __repr__(self) -> str:
        return f"WeightInfo(name='{self.name}', type='{self.matrix_type}', shape={self.shape})"


_LAYER_RE = re.compile(r"(?:blocks|layers?|encoder\.layer)\.(\d+)")


def _extract_layer_idx(name: str) -> int | None:
    """Reference extraction.py:284-290."""
    m = _LAYER_RE.search(name)
    return int(m.group(1)) if m else None


def _selected(name: str, layer_patterns: list[str] | None) -> bool:
    """Substring match, so "blocks.1" also selects blocks.10 (extraction.py:51)."""
    return not layer_patterns or any(p in name for p in layer_patterns)


def _info(name: str, layer_idx, matrix_type: str, w: torch.Tensor) -> WeightInfo:
    return WeightInfo(name=name, layer_idx=layer_idx, matrix_type=matrix_type, weight=w, shape=tuple(w.shape))


def extract_qkv_weights(model: nn.Module, layer_patterns: list[str] | None = None) -> list[WeightInfo]:
    """Fused `qkv` split into q/k/v row-blocks, or separate q_proj/k_proj/v_proj.
    Reference extraction.py:32-112."""
    out: list[WeightInfo] = []
    for name, module in model.named_modules():
        if not _selected(name, layer_patterns):
            continue
        if hasattr(module, "qkv") and hasattr(module.qkv, "weight"):
            qkv = module.qkv.weight.detach()
            d = qkv.shape[1]
            li = _extract_layer_idx(name)
            out.append(_info(f"{name}.qkv.q", li, "q", qkv[:d]))
            out.append(_info(f"{name}.qkv.k", li, "k", qkv[d : 2 * d]))
            out.append(_info(f"{name}.qkv.v", li, "v", qkv[2 * d :]))
        elif hasattr(module, "q_proj") and hasattr(module.q_proj, "weight"):
            li = _extract_layer_idx(name)
            for proj_name, proj_type in (("q_proj", "q"), ("k_proj", "k"), ("v_proj", "v")):
                proj = getattr(module, proj_name, None)
                if proj is not None and hasattr(proj, "weight"):
                    out.append(_info(f"{name}.{proj_name}", li, proj_type, proj.weight.detach()))
    return out


def extract_attention_weights(model: nn.Module, layer_patterns: list[str] | None = None) -> list[WeightInfo]:
    """Attention output projections (`.proj` under a module named *attn*/*attention*).
    Reference extraction.py:115-155."""
    out: list[WeightInfo] = []
    for name, module in model.named_modules():
        if not _selected(name, layer_patterns):
            continue
        low = name.lower()
        if hasattr(module, "proj") and hasattr(module.proj, "weight") and ("attn" in low or "attention" in low):
            out.append(_info(f"{name}.proj", _extract_layer_idx(name), "attn_proj", module.proj.weight.detach()))
    return out


def extract_mlp_weights(model: nn.Module, layer_patterns: list[str] | None = None) -> list[WeightInfo]:
    """Any module with mlp/ffn in its name and a tensor `.weight`; typed by fc1/fc2
    (or Sequential index 0/2).  Reference extraction.py:158-205."""
    out: list[WeightInfo] = []
    for name, module in model.named_modules():
        if not _selected(name, layer_patterns):
            continue
        low = name.lower()
        if ("mlp" in low or "ffn" in low) and hasattr(module, "weight") and isinstance(module.weight, torch.Tensor):
            last = name.split(".")[-1]
            if "fc1" in name or "0" in last:
                mlp_type = "mlp_up"
            elif "fc2" in name or "2" in last:
                mlp_type = "mlp_down"
            else:
                mlp_type = "mlp"
            out.append(_info(name, _extract_layer_idx(name), mlp_type, module.weight.detach()))
    return out


def extract_patch_embed_weights(model: nn.Module) -> list[WeightInfo]:
    """Conv patch embedding reshaped [out, in*h*w].  Reference extraction.py:208-242."""
    out: list[WeightInfo] = []
    for name, module in model.named_modules():
        if "patch_embed" in name.lower() and hasattr(module, "proj") and hasattr(module.proj, "weight"):
            w = module.proj.weight.detach()
            if w.ndim == 4:
                w = w.reshape(w.shape[0], -1)
            out.append(_info(f"{name}.proj", None, "patch_embed", w))
    return out


def extract_all_weights(
    model: nn.Module,
    layer_patterns: list[str] | None = None,
    include_qkv: bool = True,
    include_proj: bool = True,
    include_mlp: bool = False,
    include_patch_embed: bool = True,
) -> list[WeightInfo]:
    """Union in the reference's order: qkv, attn proj, mlp, patch embed (which
    ignores `layer_patterns`).  Reference extraction.py:245-281."""
    out: list[WeightInfo] = []
    if include_qkv:
        out.extend(extract_qkv_weights(model, layer_patterns))
    if include_proj:
        out.extend(extract_attention_weights(model, layer_patterns))
    if include_mlp:
        out.extend(extract_mlp_weights(model, layer_patterns))
    if include_patch_embed:
        out.extend(extract_patch_embed_weights(model))
    return out


def group_weights_by_layer(weights: list[WeightInfo]) -> dict[int | None, list[WeightInfo]]:
    """Reference extraction.py:293-312."""
    grouped: dict[int | None, list[WeightInfo]] = {}
    for w in weights:
        grouped.setdefault(w.layer_idx, []).append(w)
    return grouped


def group_weights_by_type(weights: list[WeightInfo]) -> dict[str, list[WeightInfo]]:
    """Reference extraction.py:315-334."""
    grouped: dict[str, list[WeightInfo]] = {}
    for w in weights:
        grouped.setdefault(w.matrix_type, []).append(w)
    return grouped
