"""CPU oracle for the SimSeg hot path.  TEST INFRASTRUCTURE ONLY.

This file restates, in plain fp32 PyTorch running on the CPU, the arithmetic of
the reference's data-parallel hot path so the CUDA kernels have something to be
checked against.  It is imported only by ``tests/``, ``__graft_entry__.smoke()``
and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.  The product
package ``simseg_b200`` never imports it.

Every function cites the reference lines it follows (paths relative to
``/root/reference``).  The ViT and BERT internals live in third-party packages
that are NOT vendored in the reference tree:

* ``timm==0.6.13`` (``requirements.txt:9``) ``VisionTransformer`` for
  ``vit_{small,base}_patch16_224_in21k`` — restated from its published
  algorithm (pre-norm blocks, LayerNorm eps 1e-6, exact-erf GELU, qkv bias,
  no LayerScale, all drop rates 0), driven token-level exactly as
  ``simseg/models/backbones/mml/vit_builder.py:13-21`` does.
* ``transformers==4.21.3`` (``requirements.txt:13``) ``BertModel`` for
  ``bert-base-uncased`` — restated (post-LN encoder, eps 1e-12, erf GELU,
  additive ``finfo.min`` key mask), called as
  ``simseg/models/backbones/mml/huggingface_builder.py:16-17``.

Pinning (see ``oracle/make_golden.py``): the reference ships no tests or golden
vectors, so the oracle is pinned against the reference's OWN modules imported
from ``/root/reference`` in the build container (heads, pooling, L2norm, NCE,
GatherLayer semantics, EmbANN, full ``CLIPModel``), against
``torchvision.models.vision_transformer.VisionTransformer`` (independent ViT
implementation) and against the installed ``transformers`` ``BertModel``.  The
outputs are committed under ``tests/golden/``.

``zero_shot_class_embedding``, ``seg_select`` and ``seg_norm_maps`` restate lines of ``tools/seg_evaluation.py``, a
script that cannot be imported here (``pydensecrf`` is absent): ``make_golden.py:seg_glue`` executes those lines read from
the reference tree (``tests/golden/seg_glue.npz``).
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------- #
# encoders
# --------------------------------------------------------------------------- #
def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float) -> Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu_erf(x: Tensor) -> Tensor:
    return 0.5 * x * (1.0 + torch.erf(x * (1.0 / math.sqrt(2.0))))


def vit_forward(sd: Dict[str, Tensor], image: Tensor, heads: int, prefix: str = "",
                depth: int = 12, eps: float = 1e-6) -> Tensor:
    """ViT token-level forward, all tokens returned.

    Follows ``simseg/models/backbones/mml/vit_builder.py:13-21`` (patch_embed ->
    cat cls -> + pos_embed -> blocks -> norm) with timm 0.6.13 block semantics
    (SURVEY.md appendix B.1).  ``sd`` uses timm state-dict names.
    """
    p = lambda k: sd[prefix + k]
    B = image.shape[0]
    w = p("patch_embed.proj.weight")                     # (D,3,16,16)
    D = w.shape[0]
    x = F.conv2d(image, w, p("patch_embed.proj.bias"), stride=16)   # (B,D,h,w)
    x = x.flatten(2).transpose(1, 2)                     # (B,N,D)
    x = torch.cat([p("cls_token").expand(B, -1, -1), x], dim=1)
    x = x + p("pos_embed")
    S = x.shape[1]
    hd = D // heads
    for i in range(depth):
        b = f"blocks.{i}."
        y = layer_norm(x, p(b + "norm1.weight"), p(b + "norm1.bias"), eps)
        qkv = y @ p(b + "attn.qkv.weight").T + p(b + "attn.qkv.bias")
        qkv = qkv.reshape(B, S, 3, heads, hd).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        a = torch.softmax((q @ k.transpose(-2, -1)) * hd ** -0.5, dim=-1)
        o = (a @ v).transpose(1, 2).reshape(B, S, D)
        x = x + (o @ p(b + "attn.proj.weight").T + p(b + "attn.proj.bias"))
        y = layer_norm(x, p(b + "norm2.weight"), p(b + "norm2.bias"), eps)
        h = gelu_erf(y @ p(b + "mlp.fc1.weight").T + p(b + "mlp.fc1.bias"))
        x = x + (h @ p(b + "mlp.fc2.weight").T + p(b + "mlp.fc2.bias"))
    return layer_norm(x, p("norm.weight"), p("norm.bias"), eps)


# ---- train-mode dropout of the BERT tower ----------------------------------------------------------------------------
# HF ``BertModel`` under ``model.train()`` (the reference trains it so: ``huggingface_builder.py:16-17`` inside
# ``CLIPModel``; bert-base-uncased config: hidden_dropout_prob = attention_probs_dropout_prob = 0.1) applies dropout
#   site 0          after the embedding LayerNorm                       (BertEmbeddings)
#   site 1 + 3 l    to the attention probabilities, after the softmax   (BertSelfAttention)
#   site 2 + 3 l    to attention.output.dense before the residual add   (BertSelfOutput)
#   site 3 + 3 l    to output.dense before the residual add             (BertOutput)
# — the order HF calls them in; ``make_golden.py:bert_dropout`` pins the placement against the installed BertModel by serving
# it these very masks.  WHICH elements drop is each framework's own random stream (torch's CUDA Philox offsets depend on its
# kernel launch geometry), so the product defines its stream as below and the tests compare against it bit for bit.
_PHILOX_M0, _PHILOX_M1 = 0xD2511F53, 0xCD9E8D57
_PHILOX_W0, _PHILOX_W1 = 0x9E3779B9, 0xBB67AE85


def philox4x32_10(ctr, key):
    """Philox4x32-10 (Salmon et al., SC'11; the Random123 reference algorithm).  ``ctr``: uint32 array [..., 4],
    ``key``: two uint32.  Returns uint32 [..., 4].  Known answers: tests/test_oracle_cpu.py::test_philox_known_answers."""
    import numpy as np
    c = [np.asarray(ctr[..., i], dtype=np.uint64) for i in range(4)]
    k0, k1 = int(key[0]) & 0xFFFFFFFF, int(key[1]) & 0xFFFFFFFF
    mask = np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0 = np.uint64(_PHILOX_M0) * c[0]
        p1 = np.uint64(_PHILOX_M1) * c[2]
        hi0, lo0 = p0 >> np.uint64(32), p0 & mask
        hi1, lo1 = p1 >> np.uint64(32), p1 & mask
        c = [hi1 ^ c[1] ^ np.uint64(k0), lo1, hi0 ^ c[3] ^ np.uint64(k1), lo0]
        k0 = (k0 + _PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + _PHILOX_W1) & 0xFFFFFFFF
    return np.stack(c, axis=-1).astype(np.uint32)


def dropout_threshold(p: float) -> int:
    """An element is dropped iff its 32-bit word < round(p * 2**32) (csrc/philox.cuh:make_drop_spec)."""
    t = float(p) * 4294967296.0
    return 0xFFFFFFFF if t >= 4294967295.0 else int(t + 0.5)


def hidden_keep_mask(seed: int, step: int, site: int, M: int, D: int, p: float) -> Tensor:
    """bool [M, D]: element (row, col) is word (col & 3) of philox(key = seed, counter = {col >> 2, row, site, step})."""
    import numpy as np
    assert D % 4 == 0
    rows, grp = np.meshgrid(np.arange(M, dtype=np.uint32), np.arange(D // 4, dtype=np.uint32), indexing="ij")
    ctr = np.stack([grp, rows, np.full_like(rows, site), np.full_like(rows, step & 0xFFFFFFFF)], axis=-1)
    w = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)).reshape(M, D)
    return torch.from_numpy(w >= np.uint32(dropout_threshold(p)))


def attn_keep_mask(seed: int, step: int, site: int, B: int, H: int, S: int, p: float) -> Tensor:
    """bool [B, H, S, S]: (b, h, q, k) is word (k & 3) of philox(key = seed, counter = {k >> 2, (b H + h) S + q, site, step})."""
    import numpy as np
    ng = (S + 3) // 4
    rows, grp = np.meshgrid(np.arange(B * H * S, dtype=np.uint32), np.arange(ng, dtype=np.uint32), indexing="ij")
    ctr = np.stack([grp, rows, np.full_like(rows, site), np.full_like(rows, step & 0xFFFFFFFF)], axis=-1)
    w = philox4x32_10(ctr, (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF)).reshape(B * H * S, ng * 4)[:, :S]
    return torch.from_numpy(w >= np.uint32(dropout_threshold(p))).reshape(B, H, S, S)


class PhiloxDropout:
    """The product's dropout stream for one forward: ``hidden(site, x)`` / ``attn(site, probs)`` return the dropped tensor."""

    def __init__(self, seed: int, step: int, p_hidden: float = 0.1, p_attn: float = 0.1):
        self.seed, self.step, self.p_hidden, self.p_attn = seed, step, p_hidden, p_attn

    def hidden(self, site: int, x: Tensor) -> Tensor:
        if self.p_hidden <= 0:
            return x
        keep = hidden_keep_mask(self.seed, self.step, site, x.numel() // x.shape[-1], x.shape[-1], self.p_hidden)
        return x * keep.reshape(x.shape).to(x.dtype) * (1.0 / (1.0 - self.p_hidden))

    def attn(self, site: int, a: Tensor) -> Tensor:
        if self.p_attn <= 0:
            return a
        B, H, S, _ = a.shape
        return a * attn_keep_mask(self.seed, self.step, site, B, H, S, self.p_attn).to(a.dtype) * (1.0 / (1.0 - self.p_attn))


class TorchDropout:
    """``F.dropout`` from torch's global generator — the CPU baseline's stand-in when only the WORK matters (bench.py)."""

    def __init__(self, p_hidden: float = 0.1, p_attn: float = 0.1):
        self.p_hidden, self.p_attn = p_hidden, p_attn

    def hidden(self, site: int, x: Tensor) -> Tensor:
        return F.dropout(x, self.p_hidden, True)

    def attn(self, site: int, a: Tensor) -> Tensor:
        return F.dropout(a, self.p_attn, True)


def bert_forward(sd: Dict[str, Tensor], input_ids: Tensor, attention_mask: Tensor,
                 heads: int = 12, prefix: str = "", depth: int = 12, eps: float = 1e-12, dropout=None) -> Tensor:
    """BERT encoder ``last_hidden_state``, HF state-dict names.  ``dropout`` = None (eval mode / p = 0) or an object with
    ``hidden(site, x)`` / ``attn(site, probs)`` (train mode, sites as listed above).

    Call site: ``simseg/models/pipelines/clip.py:220-223`` through
    ``huggingface_builder.py:16-17``; internals per SURVEY.md appendix B.2.
    """
    dh = (lambda site, x: x) if dropout is None else dropout.hidden
    da = (lambda site, x: x) if dropout is None else dropout.attn
    p = lambda k: sd[prefix + k]
    B, T = input_ids.shape
    e = (p("embeddings.word_embeddings.weight")[input_ids]
         + p("embeddings.token_type_embeddings.weight")[0]
         + p("embeddings.position_embeddings.weight")[:T])
    h = dh(0, layer_norm(e, p("embeddings.LayerNorm.weight"), p("embeddings.LayerNorm.bias"), eps))
    D = h.shape[-1]
    hd = D // heads
    bias = (1.0 - attention_mask.to(h.dtype))[:, None, None, :] * torch.finfo(h.dtype).min
    for i in range(depth):
        l = f"encoder.layer.{i}."
        lin = lambda t, name: t @ p(l + name + ".weight").T + p(l + name + ".bias")
        sh = lambda t: t.reshape(B, T, heads, hd).transpose(1, 2)
        q, k, v = sh(lin(h, "attention.self.query")), sh(lin(h, "attention.self.key")), sh(lin(h, "attention.self.value"))
        a = da(1 + 3 * i, torch.softmax(q @ k.transpose(-2, -1) / math.sqrt(hd) + bias, dim=-1))
        c = (a @ v).transpose(1, 2).reshape(B, T, D)
        h = layer_norm(dh(2 + 3 * i, lin(c, "attention.output.dense")) + h,
                       p(l + "attention.output.LayerNorm.weight"), p(l + "attention.output.LayerNorm.bias"), eps)
        f = gelu_erf(lin(h, "intermediate.dense"))
        h = layer_norm(dh(3 + 3 * i, lin(f, "output.dense")) + h,
                       p(l + "output.LayerNorm.weight"), p(l + "output.LayerNorm.bias"), eps)
    return h


# --------------------------------------------------------------------------- #
# heads
# --------------------------------------------------------------------------- #
def simple_projection(x: Tensor, w: Tensor) -> Tensor:
    """``SimpleProjection.forward`` — ``simseg/models/components/projection.py:45-46``."""
    return x @ w.T


def topk_pooling(x: Tensor, k: int, attention_mask: Optional[Tensor] = None) -> Tensor:
    """LoDA pooling — ``simseg/models/components/pooling.py:52-65``.

    Per (sample, channel) mean of the top-k values over the token axis; masked
    variant writes -10000 into padded tokens first (``pooling.py:60``) and
    shrinks k to the shortest caption (``pooling.py:61-63``).  Not in-place
    here (the reference mutates its input).
    """
    if attention_mask is not None:
        x = x.clone()
        x[attention_mask == 0] = -10000.0
        k = min(k, int(attention_mask.sum(1).min()))
    vals = x.topk(k, dim=1)[0]
    return vals.mean(1)


def l2norm(x: Tensor, eps: float = 1e-8) -> Tensor:
    """``L2norm`` — ``simseg/models/components/normalization.py:6-11`` (eps ADDED)."""
    return x / (x.pow(2).sum(-1, keepdim=True).sqrt() + eps)


def image_embed(tokens: Tensor, w_img: Tensor, k: int = 5) -> Tensor:
    """``forward_image_feature`` drop-CLS + ``forward_image_project`` — ``clip.py:65-93``."""
    return l2norm(topk_pooling(simple_projection(tokens[:, 1:], w_img), k))


def text_embed(tokens: Tensor, w_txt: Tensor, attention_mask: Tensor, k: int = 1) -> Tensor:
    """``forward_text_feature`` slice + ``forward_text_project`` — ``clip.py:96-120``."""
    return l2norm(topk_pooling(simple_projection(tokens, w_txt), k, attention_mask))


# --------------------------------------------------------------------------- #
# loss
# --------------------------------------------------------------------------- #
def nce_direction(feat1: Tensor, feat2_global: Tensor, temperature: Tensor, rank: int = 0
                  ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """One ``NCE.forward`` call, global-reduce branch — ``mml_loss.py:51-96``.

    ``feat1`` (b,E) local rows, ``feat2_global`` (W*b,E) gathered columns.
    Returns (loss scalar, acc scalar, logits (b,W*b), per-row CE (b,)).
    ignore_mask is all zeros in the shipped path (``clip.py:171-175``).
    """
    b = feat1.shape[0]
    temp = torch.clamp(temperature, 0.001, 0.5)                   # mml_loss.py:56
    logits = (feat1 @ feat2_global.T) / temp                      # :73
    targets = torch.arange(b * rank, b * (rank + 1), device=feat1.device)   # :75
    rows = F.cross_entropy(logits, targets, reduction="none")     # :77
    loss = rows.mean()                                            # :89-91
    acc = (logits.argmax(1) == targets).float().sum() / b         # utils/misc.py:462-477
    return loss, acc, logits, rows


def clip_loss(img_local: Tensor, txt_local: Tensor, img_global: Tensor, txt_global: Tensor,
              temperature: Tensor, rank: int = 0):
    """``CLIPModel.forward_loss`` global_reduce branch — ``clip.py:123-149``."""
    l_i2t, a_i2t, _, _ = nce_direction(img_local, txt_global, temperature, rank)
    l_t2i, a_t2i, _, _ = nce_direction(txt_local, img_global, temperature, rank)
    return 0.5 * (l_i2t + l_t2i), a_i2t, a_t2i


# --------------------------------------------------------------------------- #
# dense patch-text similarity map (zero-shot segmentation)
# --------------------------------------------------------------------------- #
def patch_text_sim(patch_emb: Tensor, text_emb: Tensor) -> Tuple[Tensor, Tensor]:
    """``tools/seg_evaluation.py:111-112,136`` for every image and every class.

    ``patch_emb`` (B,N,E) projected patch tokens, ``text_emb`` (C,E) class
    embeddings.  ``F.normalize`` (x / max(||x||, 1e-12)) then dot with each
    class.  Returns sim (B,N,C) fp32 and argmax (B,N) int64.
    """
    p = F.normalize(patch_emb.float(), dim=-1, p=2)
    sim = p @ text_emb.float().T
    return sim, sim.argmax(-1)


def zero_shot_class_embedding(prompt_emb: Tensor) -> Tensor:
    """``tools/seg_evaluation.py:71-72``: mean over prompts, then /= norm (no eps).
    ``prompt_emb`` (C,P,E) -> (C,E)."""
    m = prompt_emb.mean(1)
    return m / m.norm(dim=-1, keepdim=True)


def image_level_scores(img_emb: Tensor, text_emb: Tensor) -> Tensor:
    """``tools/seg_evaluation.py:119``: (B,E)·(C,E) -> (B,C)."""
    return (img_emb[:, None, :] * text_emb[None]).sum(-1)


def seg_select(img_emb: Tensor, text_emb: Tensor, top_cls_num: int, max_cand: int = 5):
    """``tools/seg_evaluation.py:119-128,141-144`` for every image of a batch: image-level scores, top-k, threshold
    ``mean + std`` (torch.std, unbiased), then the reference loop over the first ``max_cand`` classes of the top-k: ids 0
    and 255 are skipped, the scan stops at the first score below the threshold.  Returns (scores, cand padded with -1,
    threshold)."""
    scores = image_level_scores(img_emb.float(), text_emb.float())
    B = scores.shape[0]
    cand = torch.full((B, max_cand), -1, dtype=torch.int32)
    thr = torch.zeros(B)
    for b in range(B):
        tv, ti = scores[b].topk(top_cls_num)
        thr[b] = tv.mean() + 1.0 * tv.std()
        n = 0
        for i, index in enumerate(ti[:max_cand]):
            if int(index) in (0, 255):
                continue
            if float(scores[b, index]) < float(thr[b]):
                break
            cand[b, n] = int(index)
            n += 1
    return scores, cand, thr


def seg_norm_maps(sim: Tensor, cand: Tensor, h: int, w: int, scale: int = 16) -> Tensor:
    """``tools/seg_evaluation.py:131-139,146-147``: class column of the map -> (h,w) -> nearest x16 -> min-max normalisation."""
    B, N, _ = sim.shape
    K = cand.shape[1]
    out = torch.zeros(B, K, h * scale, w * scale)
    for b in range(B):
        for k in range(K):
            c = int(cand[b, k])
            if c < 0:
                continue
            a = upsample_nearest(sim[b, :, c].reshape(h, w), scale)
            out[b, k] = (a - a.min()) / (a.max() - a.min())
    return out


def upsample_nearest(sim_map: Tensor, scale: int = 16) -> Tensor:
    """``tools/seg_evaluation.py:137-139`` — (…,h,w) -> (…,h*scale,w*scale)."""
    return sim_map.repeat_interleave(scale, -2).repeat_interleave(scale, -1)


# --------------------------------------------------------------------------- #
# retrieval
# --------------------------------------------------------------------------- #
def allpairs_sim(left: Tensor, right: Tensor) -> Tensor:
    """``EmbANN._ann`` matmul — ``simseg/tasks/clip/hooks/utils.py:36``."""
    return left.float() @ right.float().T


def retrieval_first_match_rank(sim: Tensor, left_gid: Tensor, right_gid: Tensor) -> Tuple[Tensor, Tensor]:
    """``EmbANN._ann`` :37-42 + ``RetrievalMetric.__call__`` :63-65.

    Returns (has_match (M,) bool, rank of the first matching right item (M,)).
    """
    order = torch.argsort(sim, dim=1, descending=True)
    gid_sorted = right_gid[None].expand_as(sim).gather(1, order)
    matched = gid_sorted == left_gid[:, None]
    has, first = torch.max(matched, dim=1)
    return has, first


def recall_at(has: Tensor, first: Tensor, ks=(1, 5, 10)) -> Dict[str, float]:
    """``RetrievalMetric.__call__`` — ``hooks/utils.py:66-71``."""
    rank = first[has]
    return {f"R@{k}": ((rank < k).sum() / has.sum()).item() for k in ks}


# --------------------------------------------------------------------------- #
# position-embedding resize
# --------------------------------------------------------------------------- #
def interpolate_pos_embed(pos_embed: Tensor, new_num_patches: int, num_extra: int = 1) -> Tensor:
    """``simseg/utils/interpolate_pe.py:4-27`` (bicubic, align_corners=False)."""
    E = pos_embed.shape[-1]
    orig = int((pos_embed.shape[-2] - num_extra) ** 0.5)
    new = int(new_num_patches ** 0.5)
    if orig == new:
        return pos_embed
    extra, pos = pos_embed[:, :num_extra], pos_embed[:, num_extra:]
    pos = pos.reshape(-1, orig, orig, E).permute(0, 3, 1, 2)
    pos = F.interpolate(pos, size=(new, new), mode="bicubic", align_corners=False)
    return torch.cat([extra, pos.permute(0, 2, 3, 1).flatten(1, 2)], dim=1)


# --------------------------------------------------------------------------- #
# whole model, functional (state-dict keys as SURVEY.md §8b)
# --------------------------------------------------------------------------- #
IMG_PREFIX = "image_encoder.model.model."
TXT_PREFIX = "text_encoder.model.model."


def clip_embeddings(sd: Dict[str, Tensor], batch: Dict[str, Tensor], vit_heads: int,
                    image_k: int = 5, text_k: int = 1, dropout=None) -> Tuple[Tensor, Tensor]:
    """``CLIPModel.forward(batch, embeddings='all')`` — ``clip.py:152-168``.  ``dropout``: BERT train mode (``bert_forward``)."""
    it = vit_forward(sd, batch["image"], vit_heads, IMG_PREFIX)
    tt = bert_forward(sd, batch["input_ids"], batch["attention_mask"], 12, TXT_PREFIX, dropout=dropout)
    img = image_embed(it, sd["image_projection.linear.weight"], image_k)
    txt = text_embed(tt, sd["text_projection.linear.weight"], batch["attention_mask"], text_k)
    return img, txt


def clip_train_forward(sd: Dict[str, Tensor], batch: Dict[str, Tensor], vit_heads: int,
                       image_k: int = 5, text_k: int = 1, dropout=None):
    """``CLIPModel.forward(batch)`` at world size 1 — ``clip.py:152-176``."""
    img, txt = clip_embeddings(sd, batch, vit_heads, image_k, text_k, dropout)
    return clip_loss(img, txt, img, txt, sd["loss.temperature"], 0)


# --------------------------------------------------------------------------- #
# seeded synthetic weights / inputs (shared by tests, smoke and bench)
# --------------------------------------------------------------------------- #
def make_state_dict(vit_dim: int, vit_heads: int, img_size: int = 224, depth: int = 12,
                    txt_depth: int = 12, seed: int = 0, proj_dim: int = 512,
                    vocab: int = 30522, txt_dim: int = 768, txt_ffn: int = 3072,
                    max_pos: int = 512) -> Dict[str, Tensor]:
    """Random-init weights with the reference's state-dict names.

    trunc-normal(0.02)-like init (plain normal*0.02), LayerNorm 1/0 perturbed a
    little so gamma/beta paths are exercised.
    """
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, std=0.02: torch.randn(*s, generator=g) * std
    sd: Dict[str, Tensor] = {}
    D, N = vit_dim, (img_size // 16) ** 2
    P = IMG_PREFIX
    sd[P + "cls_token"] = rn(1, 1, D)
    sd[P + "pos_embed"] = rn(1, N + 1, D)
    sd[P + "patch_embed.proj.weight"] = rn(D, 3, 16, 16)
    sd[P + "patch_embed.proj.bias"] = rn(D)
    for i in range(depth):
        b = f"{P}blocks.{i}."
        for n in ("norm1", "norm2"):
            sd[b + n + ".weight"] = 1.0 + rn(D, std=0.05)
            sd[b + n + ".bias"] = rn(D, std=0.05)
        sd[b + "attn.qkv.weight"] = rn(3 * D, D); sd[b + "attn.qkv.bias"] = rn(3 * D)
        sd[b + "attn.proj.weight"] = rn(D, D); sd[b + "attn.proj.bias"] = rn(D)
        sd[b + "mlp.fc1.weight"] = rn(4 * D, D); sd[b + "mlp.fc1.bias"] = rn(4 * D)
        sd[b + "mlp.fc2.weight"] = rn(D, 4 * D); sd[b + "mlp.fc2.bias"] = rn(D)
    sd[P + "norm.weight"] = 1.0 + rn(D, std=0.05)
    sd[P + "norm.bias"] = rn(D, std=0.05)
    Q, H = TXT_PREFIX, txt_dim
    sd[Q + "embeddings.word_embeddings.weight"] = rn(vocab, H)
    sd[Q + "embeddings.position_embeddings.weight"] = rn(max_pos, H)
    sd[Q + "embeddings.token_type_embeddings.weight"] = rn(2, H)
    sd[Q + "embeddings.LayerNorm.weight"] = 1.0 + rn(H, std=0.05)
    sd[Q + "embeddings.LayerNorm.bias"] = rn(H, std=0.05)
    for i in range(txt_depth):
        l = f"{Q}encoder.layer.{i}."
        for n in ("attention.self.query", "attention.self.key", "attention.self.value", "attention.output.dense"):
            sd[l + n + ".weight"] = rn(H, H); sd[l + n + ".bias"] = rn(H)
        sd[l + "intermediate.dense.weight"] = rn(txt_ffn, H); sd[l + "intermediate.dense.bias"] = rn(txt_ffn)
        sd[l + "output.dense.weight"] = rn(H, txt_ffn); sd[l + "output.dense.bias"] = rn(H)
        for n in ("attention.output.LayerNorm", "output.LayerNorm"):
            sd[l + n + ".weight"] = 1.0 + rn(H, std=0.05)
            sd[l + n + ".bias"] = rn(H, std=0.05)
    sd["image_projection.linear.weight"] = rn(proj_dim, D, std=D ** -0.5)
    sd["text_projection.linear.weight"] = rn(proj_dim, H, std=H ** -0.5)
    sd["loss.temperature"] = torch.tensor(0.02)
    return sd


def make_batch(B: int, T: int = 25, img_size: int = 224, seed: int = 1234, vocab: int = 30522,
               min_len: int = 8) -> Dict[str, Tensor]:
    """Synthetic batch per SURVEY.md §8d: image ~ N(0,1); ids ~ U[0,vocab) with [CLS]=101 first;
    attention_mask = ones up to a length ~ U{min_len..T} then zeros."""
    g = torch.Generator().manual_seed(seed)
    image = torch.randn(B, 3, img_size, img_size, generator=g)
    ids = torch.randint(0, vocab, (B, T), generator=g)
    ids[:, 0] = 101
    lens = torch.randint(min(min_len, T), T + 1, (B,), generator=g)
    mask = (torch.arange(T)[None] < lens[:, None]).long()
    return {"image": image, "input_ids": ids, "attention_mask": mask}
