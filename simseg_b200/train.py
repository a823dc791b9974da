"""Data-parallel training step around ``CLIPModel`` (what ``CLIPRunner.batch_processor`` +
``OptimizerHook.after_train_step`` + the DDP wrap do in the reference: ``tasks/clip/clip_runner.py:216-251``,
``core/hooks/optimizer.py:69-87``, ``core/hooks/dist.py:47-51``).

* gradients of each tower live in one flat fp32 buffer (``dist.FlatGrads``); its mean all-reduce is enqueued the
  moment that tower's backward has been issued, so it overlaps the other tower's backward over NVLink;
* the optimizer is the one the YAML names (``optim.name: torch.optim.AdamW``) — host-side PyTorch, fused kernel;
* optional micro-batching with an embedding cache (the reference's BSGS idea, ``tasks/clip/clip_bsgs_runner.py:
  309-451``): pass 1 embeds micro-batches without saving activations, the loss and the embedding gradients are
  computed on the full (gathered) batch, pass 2 re-runs each micro-batch with activations and back-propagates the
  cached embedding gradient.  Mathematically identical to the single pass; trades 1 extra forward for memory.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from . import dist as sdist
from .pipeline import CLIPModel

Tensor = torch.Tensor


class Trainer:
    def __init__(self, model: CLIPModel, cfg, micro_batch: Optional[int] = None):
        self.model, self.cfg, self.micro_batch = model, cfg, micro_batch
        vit = list(model.image_encoder.parameters())
        bert = list(model.text_encoder.parameters())
        heads = [p for n, p in model.named_parameters() if not n.startswith(("image_encoder.", "text_encoder."))]
        self.flat = {"vit": sdist.FlatGrads(vit), "bert": sdist.FlatGrads(bert), "heads": sdist.FlatGrads(heads)}
        model._shared.tower_done = self._tower_done
        p = cfg.optim.param
        self.opt = torch.optim.AdamW(model.parameters(), lr=cfg.optim.lr.init, betas=tuple(p.betas), eps=p.eps,
                                     weight_decay=p.weight_decay, fused=True)
        self._defer_reduce = False

    def _tower_done(self, name: str):
        if not self._defer_reduce:
            self.flat[name].all_reduce_async()

    def zero_grad(self):
        for f in self.flat.values():
            f.zero()

    def step(self, batch: Dict[str, Tensor]):
        self.zero_grad()
        B = batch["image"].shape[0]
        if self.micro_batch and self.micro_batch < B:
            out = self._step_cached(batch)
        else:
            loss_dict, i2t, t2i = self.model(batch)
            loss = loss_dict["nce_loss"]
            loss.backward()
            out = (loss.detach(), i2t, t2i)
        self.flat["heads"].all_reduce_async()
        for f in self.flat.values():
            f.wait()
        self.opt.step()
        return out

    # ---- two-pass micro-batched step with an embedding cache ---------------------------------
    def _step_cached(self, batch):
        m, mb = self.model, self.micro_batch
        B = batch["image"].shape[0]
        chunks = [slice(i, min(i + mb, B)) for i in range(0, B, mb)]
        with torch.no_grad():
            embs = [m({k: v[c] for k, v in batch.items()}, embeddings="all") for c in chunks]
        img = torch.cat([e[0] for e in embs]).requires_grad_(True)
        txt = torch.cat([e[1] for e in embs]).requires_grad_(True)
        loss_dict, i2t, t2i = m.forward_loss(img, txt)
        loss = loss_dict["nce_loss"]
        loss.backward()                                   # -> img.grad, txt.grad, temperature.grad
        self._defer_reduce = True
        try:
            for n, c in enumerate(chunks):
                if n == len(chunks) - 1:
                    self._defer_reduce = False            # the last micro-batch completes the tower gradients
                ie, te = m({k: v[c] for k, v in batch.items()}, embeddings="all")
                torch.autograd.backward([ie, te], [img.grad[c], txt.grad[c]])
        finally:
            self._defer_reduce = False
        return loss.detach(), i2t, t2i
