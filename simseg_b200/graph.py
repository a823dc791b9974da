"""CUDA-graph replay of a forward call.

The eval tools run small batches (``tools/seg_evaluation.py:99-139``: one image per step; BASELINE configs[0] / configs[3]:
32 / 64 images) through ~300-600 kernel launches of the C ABI.  At those sizes the GPU work of a launch is shorter than the
host needs to issue it (ctypes call + tensor-map encode, ~5-10 us), so the host path — not the kernels — sets the latency.
Every launch argument of a forward is static once shapes are fixed (buffers from the graph's private pool, tensor maps
encoded on the host at capture time with those addresses), so the whole call is recorded once and replayed as ONE launch;
``train.GraphedStep`` does the same for the training step.

    seg_fn = GraphedCall(lambda image: seg.segment(model, image, class_emb, top_cls_num=3), example_image)
    sim, argmax, scores, cand, maps = seg_fn(next_image)        # static outputs, overwritten by the next call
"""
from __future__ import annotations

from typing import Callable, Sequence

import torch

Tensor = torch.Tensor


def _flatten(out):
    if isinstance(out, Tensor):
        return [out]
    if isinstance(out, (tuple, list)):
        r = []
        for o in out:
            r += _flatten(o)
        return r
    if isinstance(out, dict):
        r = []
        for o in out.values():
            r += _flatten(o)
        return r
    return []


class GraphedCall:
    """``fn(*tensors) -> tensor | tuple | dict`` recorded into a CUDA graph (under ``torch.no_grad()``).

    ``fn`` must be free of host synchronisation (``.item()``, ``.tolist()``, data-dependent Python branches) and must take
    ALL its varying data through the tensor arguments; ``warmup`` eager calls run first so that lazily created state (bf16
    weight tables, ``cudaFuncSetAttribute``) exists before recording.  ``__call__`` copies new inputs into the static input
    buffers (device-to-device, or host-to-device for pinned host tensors) and replays; the returned tensors are the graph's
    static outputs and are overwritten by the next call.  Shapes and dtypes must match the example inputs."""

    def __init__(self, fn: Callable, *example_inputs: Tensor, warmup: int = 2):
        from . import ops
        if not example_inputs or not all(isinstance(t, Tensor) for t in example_inputs):
            raise TypeError("GraphedCall needs at least one example tensor input")
        dev = next((t.device for t in example_inputs if t.is_cuda), torch.device("cuda", torch.cuda.current_device()))
        self.static_in = [torch.empty(t.shape, dtype=t.dtype, device=dev) for t in example_inputs]
        for s, t in zip(self.static_in, example_inputs):
            s.copy_(t)
        self.fn = fn
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(max(warmup, 1)):
                fn(*self.static_in)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops.launch_count()
        with torch.no_grad(), torch.cuda.graph(self.graph):
            self.out = fn(*self.static_in)
        self.launches_per_replay = ops.launch_count() - n0       # kernels of the library inside one replay
        if not _flatten(self.out):
            raise ValueError("the captured call returned no tensors")

    def __call__(self, *inputs: Tensor):
        if len(inputs) != len(self.static_in):
            raise ValueError(f"expected {len(self.static_in)} inputs, got {len(inputs)}")
        for s, t in zip(self.static_in, inputs):
            if t is s:
                continue
            if t.shape != s.shape or t.dtype != s.dtype:
                raise ValueError(f"input {tuple(t.shape)} {t.dtype} does not match the recorded {tuple(s.shape)} {s.dtype}")
            s.copy_(t, non_blocking=True)
        self.graph.replay()
        return self.out


def graphed_segment(model, class_emb: Tensor, example_image: Tensor, top_cls_num: int, max_cand: int = 5,
                    patch_size: int = 16) -> GraphedCall:
    """``seg.segment`` (the per-batch body of ``tools/seg_evaluation.py:84-150`` up to the CRF) as one graph launch."""
    from . import seg
    return GraphedCall(lambda image: seg.segment(model, image, class_emb, top_cls_num, max_cand, patch_size), example_image)


def replay_sequence(calls: Sequence[Callable[[], None]]) -> torch.cuda.CUDAGraph:
    """Record ``calls`` (closures over fixed tensors, run once eagerly first) back to back into one graph — used by the
    microbenchmarks to time a short kernel on the device without the host's launch path in the measurement."""
    dev = torch.device("cuda", torch.cuda.current_device())
    side = torch.cuda.Stream(device=dev)
    side.wait_stream(torch.cuda.current_stream(dev))
    with torch.cuda.stream(side), torch.no_grad():
        for c in calls:
            c()
    torch.cuda.current_stream(dev).wait_stream(side)
    torch.cuda.synchronize(dev)
    g = torch.cuda.CUDAGraph()
    with torch.no_grad(), torch.cuda.graph(g):
        for c in calls:
            c()
    return g
