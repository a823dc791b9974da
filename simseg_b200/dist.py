"""Data-parallel plumbing: one process per GPU, ``torch.distributed`` (NCCL over NVLink/NVSwitch).

Replaces ``simseg/utils/dist.py:323-354`` (``GatherLayer``) and the DDP wrap of ``simseg/core/hooks/dist.py:47-51``:

* forward gather of the other modality's embeddings: ONE ``all_gather_into_tensor`` into a pre-shaped
  contiguous (W*b, E) buffer (the reference builds W ``zeros_like`` tensors and ``torch.cat``s them);
* backward: ``reduce_scatter_tensor(SUM)`` of the (W*b, E) gradient — each rank lands only its own b rows
  (the reference all-reduces the whole matrix and slices, ``utils/dist.py:348-354``);
* parameter gradients: all gradients of a tower live in one flat fp32 buffer that is all-reduced (mean) with a
  single asynchronous collective as soon as that tower's backward has been enqueued, overlapping the other
  tower's backward.

Everything here is device-agnostic torch.distributed code, so the N>1 logic is testable with gloo on CPU.
"""
from __future__ import annotations

from typing import Iterable, List, Optional

import torch
import torch.distributed as dist

Tensor = torch.Tensor
WORLD = "world"      # sentinel: "gather over the default process group"


def is_initialized() -> bool:
    return dist.is_available() and dist.is_initialized()


def rank() -> int:
    return dist.get_rank() if is_initialized() else 0


def world_size() -> int:
    return dist.get_world_size() if is_initialized() else 1


def _pg(group):
    """``WORLD`` -> the default process group (None for torch.distributed); anything else is a ProcessGroup."""
    return None if group is WORLD else group


def all_gather_rows(x: Tensor, group) -> Tensor:
    """(b,E) -> (W*b,E), rank-major row order (== torch.cat(all_gather(x)), utils/dist.py:338-342).
    ``group``: None = no gather, ``WORLD`` = default process group, or a ``ProcessGroup`` (group-rank-major order)."""
    if group is None or world_size() == 1:
        return x
    pg = _pg(group)
    W = dist.get_world_size(pg)
    out = torch.empty((W * x.shape[0],) + tuple(x.shape[1:]), device=x.device, dtype=x.dtype)
    if x.is_cuda:
        dist.all_gather_into_tensor(out, x.contiguous(), group=pg)
    else:                                            # gloo has no all_gather_into_tensor
        parts = list(out.chunk(W, 0))
        dist.all_gather(parts, x.contiguous(), group=pg)
    return out


def reduce_scatter_rows(g: Tensor, r: int, b: int, group) -> Tensor:
    """Sum the (W*b,E) gradients over ranks and return this rank's b rows (GatherLayer.backward semantics);
    ``r`` is the rank inside ``group``."""
    if group is None or world_size() == 1:
        return g
    pg = _pg(group)
    if g.is_cuda:
        out = torch.empty((b,) + tuple(g.shape[1:]), device=g.device, dtype=g.dtype)
        dist.reduce_scatter_tensor(out, g.contiguous(), op=dist.ReduceOp.SUM, group=pg)
        return out
    g = g.clone()                                    # gloo: all_reduce + slice (exactly the reference's path)
    dist.all_reduce(g, op=dist.ReduceOp.SUM, group=pg)
    return g[r * b:(r + 1) * b]


class FlatGrads:
    """All ``.grad`` tensors of a parameter set as views into one flat fp32 buffer + async mean all-reduce."""

    def __init__(self, params: Iterable[torch.nn.Parameter]):
        self.params: List[torch.nn.Parameter] = [p for p in params if p.requires_grad]
        if not self.params:
            raise ValueError("FlatGrads needs at least one trainable parameter (skip frozen groups)")
        n = sum(p.numel() for p in self.params)
        dev = self.params[0].device
        self.flat = torch.zeros(n, device=dev, dtype=torch.float32)
        self._attach()
        self._work = None

    def _attach(self):
        o = 0
        for p in self.params:
            p.grad = self.flat[o:o + p.numel()].view_as(p)
            o += p.numel()

    def zero(self):
        self.flat.zero_()
        if any(p.grad is None or p.grad.data_ptr() == 0 for p in self.params):
            self._attach()
        else:
            o = 0
            for p in self.params:                    # an optimizer may have replaced .grad (set_to_none=True)
                if p.grad.data_ptr() != self.flat.data_ptr() + 4 * o:
                    self._attach()
                    break
                o += p.numel()

    def all_reduce_async(self):
        """Enqueue the mean all-reduce of the flat buffer (DDP semantics) without blocking the host."""
        if world_size() == 1:
            return
        self.flat.div_(world_size())
        self._work = dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, async_op=True)

    def wait(self):
        if self._work is not None:
            self._work.wait()
            self._work = None
