"""Minimal stand-in for ``timm==0.6.13`` so the reference's ``simseg.models`` imports and
its ``ViTModel`` (``simseg/models/backbones/mml/vit_builder.py:6-21``) can be built in a
container where timm is not installed.  TEST INFRASTRUCTURE ONLY (used by
``oracle/make_golden.py``).  It restates the published architecture of timm's
``VisionTransformer`` for the two tags the shipped YAMLs name, exposing exactly the
attributes the reference touches: ``patch_embed(.num_patches,.proj)``, ``cls_token``,
``pos_embed``, ``pos_drop``, ``blocks``, ``norm`` — with timm's state-dict key names.
"""
import torch
import torch.nn as nn

__version__ = "0.6.13-shim"

_TAGS = {
    "vit_small_patch16_224_in21k": dict(dim=384, heads=6),
    "vit_base_patch16_224_in21k": dict(dim=768, heads=12),
}


class PatchEmbed(nn.Module):
    def __init__(self, img_size, dim):
        super().__init__()
        self.img_size = (img_size, img_size)
        self.num_patches = (img_size // 16) ** 2
        self.proj = nn.Conv2d(3, dim, kernel_size=16, stride=16)

    def forward(self, x):
        assert x.shape[-2:] == self.img_size
        return self.proj(x).flatten(2).transpose(1, 2)


class Attention(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.num_heads = heads
        self.scale = (dim // heads) ** -0.5
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x):
        B, S, D = x.shape
        qkv = self.qkv(x).reshape(B, S, 3, self.num_heads, D // self.num_heads).permute(2, 0, 3, 1, 4)
        q, k, v = qkv.unbind(0)
        a = ((q @ k.transpose(-2, -1)) * self.scale).softmax(dim=-1)
        return self.proj((a @ v).transpose(1, 2).reshape(B, S, D))


class Mlp(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.fc1 = nn.Linear(dim, 4 * dim)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(4 * dim, dim)

    def forward(self, x):
        return self.fc2(self.act(self.fc1(x)))


class Block(nn.Module):
    def __init__(self, dim, heads):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = Attention(dim, heads)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = Mlp(dim)

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        return x + self.mlp(self.norm2(x))


class VisionTransformer(nn.Module):
    def __init__(self, dim, heads, img_size=224, depth=12):
        super().__init__()
        self.num_heads = heads
        self.patch_embed = PatchEmbed(img_size, dim)
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.randn(1, self.patch_embed.num_patches + 1, dim) * 0.02)
        self.pos_drop = nn.Dropout(p=0.0)
        self.blocks = nn.Sequential(*[Block(dim, heads) for _ in range(depth)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)


def create_model(tag, pretrained=False, num_classes=0, img_size=224, **kwargs):
    if pretrained:
        raise RuntimeError("timm shim: no pretrained weights offline")
    cfg = _TAGS[tag]
    return VisionTransformer(cfg["dim"], cfg["heads"], img_size=img_size)
