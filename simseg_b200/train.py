"""Data-parallel training step around ``CLIPModel`` (what ``CLIPRunner.batch_processor`` +
``OptimizerHook.after_train_step`` + the DDP wrap do in the reference: ``tasks/clip/clip_runner.py:216-251``,
``core/hooks/optimizer.py:69-87``, ``core/hooks/dist.py:47-51``).

* gradients of each tower live in one flat fp32 buffer (``dist.FlatGrads``); its mean all-reduce is enqueued the
  moment that tower's backward has been issued, so it overlaps the other tower's backward over NVLink;
* the optimizer is the one the YAML names (``optim.name``, e.g. ``torch.optim.AdamW``) with the reference's grouping
  rules (``tasks/clip/hooks/optimizer.py:18-36``: base lr / weight decay, per-regex overrides from
  ``optim.param_group_rules``; groups with equal hyper-parameters are merged so the fused kernel sees few groups) and
  ``optim.grad_clip`` (``core/hooks/optimizer.py:40-49``).  LR schedules stay with the caller (control plane);
* optional micro-batching with an embedding cache (the reference's BSGS idea, ``tasks/clip/clip_bsgs_runner.py:
  309-451``): pass 1 embeds micro-batches without saving activations, the loss and the embedding gradients are
  computed on the full (gathered) batch, pass 2 re-runs each micro-batch with activations and back-propagates the
  cached embedding gradient.  Mathematically identical to the single pass at dropout p = 0 (the text top-k clamp of
  ``pooling.py:61-63`` is taken over the whole batch, as the single pass does); with BERT's train-mode dropout the two
  passes draw the SAME masks (the {seed, step} pair is restored in between — the reference's BSGS re-draws them, so its
  pass-2 activations do not match the embeddings its loss saw); trades 1 extra forward for memory.
"""
from __future__ import annotations

import importlib
import re
from typing import Dict, Optional

import torch

from . import dist as sdist
from . import towers
from .pipeline import CLIPModel

Tensor = torch.Tensor


def grouped_parameters(model: torch.nn.Module, cfg):
    """``ClipOptimizerHook.get_optimizer_grouped_parameters`` (``tasks/clip/hooks/optimizer.py:18-36``); parameters
    whose group settings are equal share one group."""
    base = {"lr": cfg.optim.lr.init, "weight_decay": cfg.optim.param["weight_decay"]}
    rules = list((cfg.optim.get("param_group_rules") or {}).values())
    merged: Dict[tuple, dict] = {}
    for key, p in model.named_parameters():
        if not p.requires_grad:
            continue
        g = dict(base)
        for rule in rules:
            if re.search(rule["regex"], key):
                g.update(rule.get("param", {}))
        k = tuple(sorted(g.items()))
        merged.setdefault(k, {**g, "params": []})["params"].append(p)
    return list(merged.values())


def build_optimizer(model: torch.nn.Module, cfg, capturable: bool = False) -> torch.optim.Optimizer:
    """``optim.name`` is a dotted class path (``core/hooks/optimizer.py:21-28`` evaluates it the same way).
    ``capturable``: optimizer state (step counters, learning rates) lives on the device so ``step()`` can be recorded
    into a CUDA graph; the learning rate of every group becomes a device scalar a scheduler may update in place."""
    mod, _, cls = str(cfg.optim.name).rpartition(".")
    ctor = getattr(importlib.import_module(mod or "torch.optim"), cls)
    kw = {k: (tuple(v) if isinstance(v, list) else v) for k, v in dict(cfg.optim.param).items()}
    kw.pop("weight_decay", None)                      # carried per group
    if ctor in (torch.optim.AdamW, torch.optim.Adam, torch.optim.SGD):
        kw.setdefault("fused", True)
    groups = grouped_parameters(model, cfg)
    if capturable:
        if ctor not in (torch.optim.AdamW, torch.optim.Adam):
            raise NotImplementedError("CUDA-graph capture of the step is wired for Adam/AdamW (capturable=True)")
        kw["capturable"] = True
        dev = groups[0]["params"][0].device
        for g in groups:
            g["lr"] = torch.tensor(float(g["lr"]), device=dev, dtype=torch.float32)
    return ctor(groups, **kw)


class Trainer:
    def __init__(self, model: CLIPModel, cfg, micro_batch: Optional[int] = None, capturable: bool = False):
        self.model, self.cfg, self.micro_batch = model, cfg, micro_batch
        vit = list(model.image_encoder.parameters())
        bert = towers.qkv_adjacent_order(model.text_encoder.named_parameters())   # q/k/v gradients of a layer back to back
        heads = [p for n, p in model.named_parameters() if not n.startswith(("image_encoder.", "text_encoder."))]
        # a frozen tower (image_encoder.trainable / text_encoder.trainable = False) simply has no buffer
        self.flat = {name: sdist.FlatGrads(ps) for name, ps in (("vit", vit), ("bert", bert), ("heads", heads))
                     if any(p.requires_grad for p in ps)}
        model._shared.tower_done = self._tower_done
        model._shared.direct_grads = True             # wgrad kernels accumulate straight into the flat buffers
        self.opt = build_optimizer(model, cfg, capturable)
        self.capturable = capturable
        gc = cfg.optim.get("grad_clip") or {}
        self.grad_clip = dict(gc) if len(gc) else None
        self._defer_reduce = False

    def _tower_done(self, name: str):
        if not self._defer_reduce and name in self.flat:
            self.flat[name].all_reduce_async()

    def zero_grad(self):
        for f in self.flat.values():
            f.zero()

    def backward_only(self, batch: Dict[str, Tensor]):
        """Forward + backward + gradient all-reduce, no optimizer step (tests, ``bench.py`` ``dp_check``)."""
        self.zero_grad()
        B = batch["image"].shape[0]
        if self.micro_batch and self.micro_batch < B:
            out = self._step_cached(batch)
        else:
            loss_dict, i2t, t2i = self.model(batch)
            loss = loss_dict["nce_loss"]
            loss.backward()
            out = (loss.detach(), i2t, t2i)
        if "heads" in self.flat:
            self.flat["heads"].all_reduce_async()
        for f in self.flat.values():
            f.wait()
        return out

    def step(self, batch: Dict[str, Tensor]):
        out = self.backward_only(batch)
        if self.grad_clip is not None:
            torch.nn.utils.clip_grad_norm_(self.model.parameters(), **self.grad_clip)
        self.opt.step()
        self.model.new_step()                         # the bf16 weight copies are re-cast (one launch) on next use
        return out

    def capture(self, batch: Dict[str, Tensor], warmup: int = 2) -> "GraphedStep":
        """Record one whole training step (forward, backward, gradient all-reduces, optimizer, weight re-cast) into a CUDA
        graph; see ``GraphedStep``."""
        return GraphedStep(self, batch, warmup)

    # ---- two-pass micro-batched step with an embedding cache ---------------------------------
    def _step_cached(self, batch):
        m, mb = self.model, self.micro_batch
        B = batch["image"].shape[0]
        chunks = [slice(i, min(i + mb, B)) for i in range(0, B, mb)]
        # pooling.py:61-63 clamps k to the shortest caption of the BATCH: fix it once for all micro-batches
        tk = m.text_pool.k
        if tk > 1:
            tk = min(tk, int(batch["attention_mask"].sum(1).min()))

        def embed(c):
            sub = {k: v[c] for k, v in batch.items()}
            ie = m.forward_image_project(m.forward_image_feature(sub["image"]))
            te = m.forward_text_project(m.forward_text_feature(sub["input_ids"], sub["attention_mask"]),
                                        sub["attention_mask"], k=tk)
            return ie, te
        hf = m.text_encoder.model
        rng0 = hf.dropout_state().clone()                 # BERT dropout: pass 2 must re-draw the masks of pass 1
        with torch.no_grad():
            embs = [embed(c) for c in chunks]
        img = torch.cat([e[0] for e in embs]).requires_grad_(True)
        txt = torch.cat([e[1] for e in embs]).requires_grad_(True)
        loss_dict, i2t, t2i = m.forward_loss(img, txt)
        loss = loss_dict["nce_loss"]
        loss.backward()                                   # -> img.grad, txt.grad, temperature.grad
        hf.drop_rng.copy_(rng0)                           # micro-batch n sees the same {seed, step} in both passes
        self._defer_reduce = True
        try:
            for n, c in enumerate(chunks):
                if n == len(chunks) - 1:
                    self._defer_reduce = False            # the last micro-batch completes the tower gradients
                ie, te = embed(c)
                torch.autograd.backward([ie, te], [img.grad[c], txt.grad[c]])
        finally:
            self._defer_reduce = False
        return loss.detach(), i2t, t2i


class GraphedStep:
    """One training step as a single CUDA-graph launch.

    A step is ~520 kernel launches through the C ABI plus ~300 small torch kernels; at the per-GPU batch of an 8-GPU run
    (512 pairs: ~35 ms of GPU work) the host needs ~15 ms to issue them and the GPU idles in between.  Everything in the step
    is static — shapes, buffers (the graph's private memory pool), tensor maps (encoded on the host at capture time with the
    pool's addresses), collectives (NCCL kernels are captured like any other) — so it is recorded once and replayed.

    ``warmup`` real steps run first on the capture stream (they DO train: lazily created buffers, bf16 weight tables and
    ``cudaFuncSetAttribute`` calls must exist before recording).  ``__call__(batch)`` copies the batch into the static input
    buffers (device-to-device, or host-to-device for a pinned host batch) and launches the graph; it returns the static
    ``(loss, i2t_acc, t2i_acc)`` tensors, overwritten by the next call."""

    def __init__(self, trainer: Trainer, batch: Dict[str, Tensor], warmup: int = 2):
        if not trainer.capturable:
            raise ValueError("build the Trainer with capturable=True to record its step into a CUDA graph")
        from . import ops
        self.trainer = trainer
        dev = next(trainer.model.parameters()).device
        self.static = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in batch.items()}
        for k, v in batch.items():
            self.static[k].copy_(v)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            for _ in range(max(warmup, 1)):
                trainer.step(self.static)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        torch.cuda.empty_cache()                      # hand the warm-up's blocks back so the graph pool can have them
        self.graph = torch.cuda.CUDAGraph()
        n0 = ops.launch_count()
        with torch.cuda.graph(self.graph):
            self.out = trainer.step(self.static)
        self.launches_per_replay = ops.launch_count() - n0      # kernels of the library inside one replay

    def __call__(self, batch: Optional[Dict[str, Tensor]] = None):
        if batch is not None and batch is not self.static:
            for k, v in batch.items():
                self.static[k].copy_(v, non_blocking=True)
        self.graph.replay()
        return self.out
