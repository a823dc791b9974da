"""Retrieval evaluation on the all-pairs similarity (BASELINE configs[4]) — host mirror of
``simseg/tasks/clip/hooks/utils.py:8-75`` (``IndexedEmbInfo``, ``EmbANN``, ``RetrievalMetric``) with the same names and
result keys.  The reference materialises ``left @ right.T`` (500 MB at 5k x 25k), argsorts every row (1 GB of int64
indices, ``utils.py:39``) and gathers group ids; here ``RetrievalMetric`` calls ONE fused tensor-core routine
(``simseg_retrieval_rank_fused``) that returns the rank of the first matching item per row and materialises neither.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Dict

import torch

from . import ops


@dataclass
class IndexedEmbInfo:
    emb_name: str
    group_idx: torch.Tensor   # [N] int64
    emb_mat: torch.Tensor     # [N,D]

    def unique(self) -> "IndexedEmbInfo":
        """One embedding per group id, ids ascending (``utils.py:14-19``: sort by id, keep the LAST row of each run).
        The reference's ``torch.sort`` is not stable, so which duplicate is "last" is unspecified there; here the sort is
        stable, i.e. the duplicate with the highest original index is kept (what the reference does on the CPU)."""
        gidx, order = torch.sort(self.group_idx, stable=True)
        uni_idx, uni_count = torch.unique_consecutive(gidx, return_counts=True)
        last = torch.cumsum(uni_count, 0) - 1
        return IndexedEmbInfo(self.emb_name, uni_idx, self.emb_mat[order[last]])

    def to_chunks(self, chunk_size):
        for start in range(0, self.emb_mat.shape[0], chunk_size):
            yield IndexedEmbInfo(self.emb_name, self.group_idx[start:start + chunk_size], self.emb_mat[start:start + chunk_size])


def first_match_rank(leftemb: IndexedEmbInfo, rightemb: IndexedEmbInfo) -> torch.Tensor:
    """int32 [M]: position of the first right item sharing the row's group id in the row's descending similarity order
    (== ``torch.max(rightgid_matched, dim=1)[1]`` of ``utils.py:63-64``; -1 where the reference's ``hasmatch`` is False)."""
    left = leftemb.emb_mat.float().contiguous()
    right = rightemb.emb_mat.float().contiguous()
    return ops.retrieval_rank_fused(left, right, leftemb.group_idx.to(torch.int64), rightemb.group_idx.to(torch.int64))


class RetrievalMetric:
    """``utils.py:52-75``: R@1 / R@5 / R@10 of left -> right retrieval, same result keys."""

    def __init__(self, with_prefix: bool = True) -> None:
        self.recall_range = (1, 5, 10)
        self.with_prefix = with_prefix

    def __call__(self, leftemb: IndexedEmbInfo, rightemb: IndexedEmbInfo) -> Dict[str, Any]:
        rank = first_match_rank(leftemb, rightemb)
        has = rank >= 0
        total = has.sum()
        assert int(total) > 0
        result: Dict[str, Any] = {}
        for bound in self.recall_range:
            result[f"R@{bound}"] = ((has & (rank < bound)).sum() / total).item()
        if self.with_prefix:
            prefix = f"[{leftemb.emb_name}] to [{rightemb.emb_name}]:"
            result = {f"{prefix} {k}": v for k, v in result.items()}
        return result
