"""Zero-shot segmentation around the patch-text map, in the call order of the reference tool
(``tools/seg_evaluation.py:57-75`` ``zero_shot_classifier``; ``:84-150`` ``evaluate_benchmark``) — every step up to the CRF
stays on the GPU (the reference fetches one class map at a time to the host).  CRF / dilate / erode post-processing is CPU
code outside the hot path (out of scope)."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops

Tensor = torch.Tensor


@torch.no_grad()
def zero_shot_classifier(model, input_ids: Tensor, attention_mask: Tensor, chunk: int = 4096) -> Tensor:
    """``input_ids`` / ``attention_mask`` (C, P, T) — P prompt captions per class (the reference tokenises 80 templates per
    class, ``utils/prompt.py``) -> unit-norm class embeddings (C, 512).  All C*P captions go through the text tower in
    batches of ``chunk`` instead of one 80-caption batch per class.  With ``text_k > 1`` the pooling clamps k to the
    shortest caption of the batch it sees (``pooling.py:61-63``) and the reference's batch is ONE class, so the clamp is
    taken per class here as well (classes with equal k share a launch)."""
    Cn, P, T = input_ids.shape
    ids, am = input_ids.reshape(Cn * P, T), attention_mask.reshape(Cn * P, T)
    k0 = model.text_pool.k
    if k0 > 1:
        kc = torch.clamp(attention_mask.sum(-1).amin(1), max=k0).tolist()          # per-class k (one host read)
    else:
        kc = [k0] * Cn
    out = torch.empty((Cn * P, model.text_projection.projection_dim), device=input_ids.device, dtype=torch.float32)
    for k in sorted(set(kc)):
        rows = torch.tensor([c for c in range(Cn) if kc[c] == k], device=ids.device)
        sel = (rows[:, None] * P + torch.arange(P, device=ids.device)[None]).reshape(-1)
        for i in range(0, sel.numel(), chunk):
            j = sel[i:i + chunk]
            feat = model.forward_text_feature(ids[j], am[j])
            out[j] = model.forward_text_project(feat, am[j], k=k).float()
    return ops.seg_class_embed(out.view(Cn, P, -1))


@torch.no_grad()
def segment(model, image: Tensor, class_emb: Tensor, top_cls_num: int, max_cand: int = 5, patch_size: int = 16):
    """One pass of ``evaluate_benchmark`` for a batch of images: returns
    ``sim`` (B, N, C) cosine map, ``argmax`` (B, N), ``scores`` (B, C) image-level class scores, ``cand`` (B, max_cand) the
    classes the reference would send to the CRF (-1 padded) and ``maps`` (B, max_cand, H, W) their min-max normalised,
    nearest-up-sampled attention maps."""
    feat = model.forward_image_feature(image)                       # (B, N, D)
    pooled = model.forward_image_project(feat)                      # (B, 512)  L2-normalised
    proj = model.image_projection(feat)                             # (B, N, 512)
    B, N, _ = proj.shape
    hw = int(round(N ** 0.5))
    text = class_emb.to(proj.dtype) if proj.dtype == torch.bfloat16 else class_emb
    sim, am = ops.patch_text_sim(proj.contiguous(), text.contiguous())
    scores, cand, _ = ops.seg_select(pooled.float(), class_emb.float(), top_cls_num, max_cand)
    maps = ops.seg_upsample_norm(sim.view(B, N, -1), cand, hw, hw, patch_size)
    return sim.view(B, N, -1), am.view(B, N), scores, cand, maps


def interpolate_pos_embed(pos_embed_checkpoint: Tensor, visual_encoder) -> Tensor:
    """Same call as the reference's ``simseg.utils.interpolate_pe.interpolate_pos_embed`` (``utils/interpolate_pe.py:4-27``;
    used at checkpoint load, ``tools/seg_evaluation.py:228-230``): resize a checkpoint's position embedding to the grid of
    ``visual_encoder`` (``.patch_embed.num_patches``, ``.pos_embed``).  Returns the input unchanged when the grids agree.
    The checkpoint tensor may live on the host (as ``torch.load(map_location='cpu')`` leaves it); the resize itself runs
    on the encoder's device and the result comes back on the checkpoint tensor's device."""
    num_patches = visual_encoder.patch_embed.num_patches
    num_extra = visual_encoder.pos_embed.shape[-2] - num_patches
    orig = int((pos_embed_checkpoint.shape[-2] - num_extra) ** 0.5)
    new = int(num_patches ** 0.5)
    if orig == new:
        return pos_embed_checkpoint
    dev = visual_encoder.pos_embed.device
    out = ops.pos_embed_bicubic(pos_embed_checkpoint.to(device=dev, dtype=torch.float32), new, num_extra)
    print("reshape position embedding from %d to %d" % (orig ** 2, new ** 2))
    return out.to(pos_embed_checkpoint.device)
