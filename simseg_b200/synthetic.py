"""Seeded synthetic inputs of the shapes BASELINE.json names: image ~ N(0,1) (B,3,224,224) fp32;
input_ids ~ U[0,30522) int64 with [CLS]=101 first; attention_mask = ones up to a length ~ U{8..T}, then zeros."""
from __future__ import annotations

from typing import Dict

import torch


def make_batch(B: int, T: int = 25, img_size: int = 224, seed: int = 1234, vocab: int = 30522,
               min_len: int = 8) -> Dict[str, torch.Tensor]:
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(B, 3, img_size, img_size, generator=g)
    ids = torch.randint(0, vocab, (B, T), generator=g)
    ids[:, 0] = 101
    lens = torch.randint(min(min_len, T), T + 1, (B,), generator=g)
    mask = (torch.arange(T)[None] < lens[:, None]).long()
    return {"image": image, "input_ids": ids, "attention_mask": mask}
