"""A timm-free stand-in for a timm VisionTransformer.

The reference's extraction is duck-typed on module names (`blocks.N.attn.qkv`,
`blocks.N.attn.proj`, `blocks.N.mlp.fc1/fc2`, `patch_embed.proj`;
extraction.py:49-62,131-153,174-203,220-240), so a model that only reproduces
those names and shapes exercises the same code.  timm is not installed here.
"""

from __future__ import annotations

import torch
import torch.nn as nn


class _Attn(nn.Module):
    def __init__(self, d):
        super().__init__()
        self.qkv = nn.Linear(d, 3 * d)
        self.proj = nn.Linear(d, d)


class _SepAttn(nn.Module):
    """q_proj/k_proj/v_proj variant (extraction.py:94-110)."""

    def __init__(self, d):
        super().__init__()
        self.q_proj = nn.Linear(d, d)
        self.k_proj = nn.Linear(d, d)
        self.v_proj = nn.Linear(d, d)
        self.proj = nn.Linear(d, d)


class _Mlp(nn.Module):
    def __init__(self, d, ratio=4):
        super().__init__()
        self.fc1 = nn.Linear(d, ratio * d)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(ratio * d, d)


class _Block(nn.Module):
    def __init__(self, d, separate_qkv=False):
        super().__init__()
        self.norm1 = nn.LayerNorm(d)
        self.attn = _SepAttn(d) if separate_qkv else _Attn(d)
        self.norm2 = nn.LayerNorm(d)
        self.mlp = _Mlp(d)


class _PatchEmbed(nn.Module):
    def __init__(self, d, in_chans=3, patch=16):
        super().__init__()
        self.proj = nn.Conv2d(in_chans, d, kernel_size=patch, stride=patch)


class StubViT(nn.Module):
    def __init__(self, embed_dim=32, depth=1, in_chans=3, patch=16, num_classes=10, seed=0, separate_qkv=False):
        super().__init__()
        self.patch_embed = _PatchEmbed(embed_dim, in_chans, patch)
        self.blocks = nn.Sequential(*[_Block(embed_dim, separate_qkv) for _ in range(depth)])
        self.norm = nn.LayerNorm(embed_dim)
        self.head = nn.Linear(embed_dim, num_classes)
        g = torch.Generator().manual_seed(seed)
        with torch.no_grad():
            for p in self.parameters():
                if p.ndim >= 2:
                    p.copy_(torch.randn(p.shape, generator=g) * 0.02)
                else:
                    p.zero_()


class WrappedViT(nn.Module):
    """Mirrors the reference's ViTClassifier, which holds the timm model under
    `.encoder` (models/vit.py:77) so names gain an `encoder.` prefix."""

    def __init__(self, **kw):
        super().__init__()
        self.encoder = StubViT(**kw)
