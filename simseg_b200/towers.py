"""Hand-sequenced forward / backward of the two towers over the C-ABI kernels.

Precision contract (= the reference under bf16 autocast, ``simseg/tasks/clip/clip_runner.py:226-228``):
fp32 master weights, bf16 tensor-core GEMM operands with fp32 accumulation, fp32 residual stream,
fp32 LayerNorm / softmax statistics.  LayerNorm outputs and MLP pre-activations are stored for backward (bf16);
GELU outputs are kept when they are a small part of the device's memory and otherwise re-emitted by the dGELU epilogue of the
fc2 dgrad GEMM (``keep_gelu_output``); what is stored per block is listed in ``_Blk``.

Gradients are accumulated (fp32) by the wgrad GEMMs into the buffer ``grad_of(param)`` names: ``param.grad`` itself
under ``train.Trainer`` (flat buffers), a scratch buffer handed back to autograd otherwise (``pipeline._VitFn``).
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional

import torch

from . import ops
from ._lib import EPI_BIAS_GELU, EPI_DGELU, EPI_NONE

Tensor = torch.Tensor


def keep_gelu_output(extra_bytes: int, device) -> bool:
    """Backward needs BOTH the fc1 pre-activation h (for gelu'(h)) and gelu(h) (operand of the fc2 wgrad).  Saving both costs
    ``extra_bytes`` (layers x tokens x 4D x 2 B); saving only h makes the dGELU epilogue of the fc2 dgrad GEMM re-emit
    gelu(h) — one more bf16 [M,4D] write per layer on a kernel that is HBM-bound (ViT-S at b = 4096: 8.0 GB -> 5.6 GB per
    launch without it).  Keep gelu(h) whenever that is a small part of the device: up to 12 % of its memory per tower
    (ViT-S: per-GPU batch <= 2900, i.e. every shard of the global batch 4096 on >= 2 GPUs; at b = 4096 on ONE GPU the step
    already peaks at 147 of 180 GB and re-emits).  ``SIMSEG_KEEP_GELU=0|1`` overrides."""
    import os
    f = os.environ.get("SIMSEG_KEEP_GELU")
    if f is not None and f != "auto":
        return f not in ("0", "")
    return extra_bytes <= 0.12 * torch.cuda.get_device_properties(device).total_memory


def param_grad(p: Tensor) -> Tensor:
    """fp32 gradient buffer of a parameter (allocated zeroed on first touch)."""
    if p.grad is None:
        p.grad = torch.zeros_like(p, memory_format=torch.contiguous_format)
    return p.grad


_grad_of = param_grad


class ScratchGrads:
    """Gradient sink for the autograd-visible path: one zeroed flat fp32 buffer with a view per trainable parameter.
    The tower backward accumulates into the views; the autograd node returns them, so ``AccumulateGrad`` (and with it
    torch DDP's bucket hooks, ``simseg/core/hooks/dist.py:48-51``) sees every parameter gradient."""

    def __init__(self, params):
        self.params = list(params)
        n = sum(p.numel() for p in self.params)
        self.flat = torch.zeros(n, device=self.params[0].device, dtype=torch.float32)
        self.views, o = {}, 0
        for p in self.params:
            self.views[id(p)] = self.flat[o:o + p.numel()].view(p.shape)
            o += p.numel()

    def __call__(self, p: Tensor) -> Tensor:
        return self.views[id(p)]

    def as_tuple(self):
        """One entry per parameter, in order; ``None`` for frozen ones."""
        return tuple(self.views[id(p)] if p.requires_grad else None for p in self.params)


class Bf16Weights:
    """bf16 copies of the fp32 master weights, refreshed once per optimizer step by ONE multi-tensor cast launch.

    The destination buffers (plain [N,K], transposed [K,N] for dgrad, and the row-packed q/k/v matrices of BERT) are
    allocated when a weight is first requested and then live as long as the model; ``clear()`` only marks them stale.
    The next request re-casts every registered weight with a single ``simseg_cast_bf16_multi`` launch (the item table is
    a device array rebuilt only when a new weight/buffer was registered or a parameter's storage moved)."""

    def __init__(self):
        self._items: Dict[tuple, dict] = {}        # (id(param), slot) -> item
        self._views: Dict[object, Tensor] = {}     # what get/get_t/packed/packed_t hand out
        self._stale = False
        self._dirty = True                         # table must be rebuilt
        self._table = None
        self._nblocks = 0
        self._keep = []

    # ---- registration ------------------------------------------------------------------------
    def _item(self, p: Tensor, slot) -> dict:
        k = (id(p), slot)
        it = self._items.get(k)
        if it is None:
            w = p.detach()
            rows = w.shape[0]
            it = {"p": p, "rows": rows, "cols": w.numel() // rows, "dst": None, "dst_t": None, "ld": 0, "ld_t": 0, "ptr": 0}
            self._items[k] = it
            self._dirty = True
        return it

    def _fill_now(self, it: dict):
        """First use inside a step: cast just this weight (later steps go through the batched refresh)."""
        w = it["p"].detach().contiguous()
        lib = ops._lib.load()
        if it["ld"] in (0, it["cols"]) and it["ld_t"] in (0, it["rows"]):
            ops.check(lib.simseg_cast_bf16(ops.ctx(), ops._p(w), ops._p(it["dst"]), ops._p(it["dst_t"]), it["rows"], it["cols"],
                                           ops._stream()), "cast_bf16")
        else:                                       # strided destination (packed matrices): use the table path for one item
            self._launch([it])

    def _launch(self, items):
        import ctypes as C
        n = len(items)
        host = torch.empty((n, 8), dtype=torch.int64)
        fb = 0
        for i, it in enumerate(items):
            w = it["p"].detach()
            assert w.is_contiguous() and w.dtype == torch.float32
            it["ptr"] = w.data_ptr()
            host[i, 0] = w.data_ptr()
            host[i, 1] = it["dst"].data_ptr() if it["dst"] is not None else 0
            host[i, 2] = it["dst_t"].data_ptr() if it["dst_t"] is not None else 0
            host[i, 3], host[i, 4] = it["rows"], it["cols"]
            host[i, 5] = it["ld"] or it["cols"]
            host[i, 6] = it["ld_t"] or it["rows"]
            host[i, 7] = fb
            fb += ((it["rows"] + 31) // 32) * ((it["cols"] + 127) // 128)     # one block = 32 rows x 128 columns
        dev = next(iter(items))["p"].device
        table = host.to(dev)
        self._keep = [table]
        ops.check(ops._lib.load().simseg_cast_bf16_multi(ops.ctx(), ops._p(table), n, fb, ops._stream()), "cast_bf16_multi")
        return table, fb

    def _refresh(self):
        if not self._stale:
            return
        items = [it for it in self._items.values() if it["dst"] is not None or it["dst_t"] is not None]
        if items:
            moved = any(it["ptr"] != it["p"].detach().data_ptr() for it in items)
            if self._dirty or moved or self._table is None:
                self._table, self._nblocks = self._launch(items)
                self._dirty = False
            else:
                ops.check(ops._lib.load().simseg_cast_bf16_multi(ops.ctx(), ops._p(self._table), len(items), self._nblocks,
                                                                ops._stream()), "cast_bf16_multi")
        self._stale = False

    # ---- accessors ---------------------------------------------------------------------------
    def get(self, p: Tensor) -> Tensor:
        self._refresh()
        it = self._item(p, "w")
        if it["dst"] is None:
            it["dst"] = torch.empty((it["rows"], it["cols"]), device=p.device, dtype=torch.bfloat16)
            it["ld"] = it["cols"]
            self._dirty = True
            self._fill_now({**it, "dst_t": None, "ld_t": 0})
        return it["dst"]

    def get_t(self, p: Tensor) -> Tensor:
        """bf16 W^T [K_in, N_out]: the K-major B operand of dgrad (dx = dy @ W), written by the same cast kernel."""
        self._refresh()
        it = self._item(p, "w")
        if it["dst_t"] is None:
            it["dst_t"] = torch.empty((it["cols"], it["rows"]), device=p.device, dtype=torch.bfloat16)
            it["ld_t"] = it["rows"]
            self._dirty = True
            self._fill_now({**it, "dst": None, "ld": 0})
        return it["dst_t"]

    def packed(self, key: str, ps: List[Tensor]) -> Tensor:
        """Row-concatenation of several [n_i, K] weights as one bf16 matrix (BERT q/k/v -> one GEMM)."""
        self._refresh()
        buf = self._views.get(key)
        if buf is None:
            K = ps[0].shape[1]
            buf = torch.empty((sum(p.shape[0] for p in ps), K), device=ps[0].device, dtype=torch.bfloat16)
            self._views[key] = buf
            r = 0
            for p in ps:
                it = self._item(p, ("pk", key))
                it["dst"], it["ld"] = buf[r:r + p.shape[0]], K
                r += p.shape[0]
                self._fill_now({**it, "dst_t": None, "ld_t": 0})
            self._dirty = True
        return buf

    def packed_t(self, key: str, ps: List[Tensor]) -> Tensor:
        """Transpose of ``packed``: [K, sum n_i]."""
        self._refresh()
        k = ("t", key)
        buf = self._views.get(k)
        if buf is None:
            K = ps[0].shape[1]
            tot = sum(p.shape[0] for p in ps)
            buf = torch.empty((K, tot), device=ps[0].device, dtype=torch.bfloat16)
            self._views[k] = buf
            c = 0
            for p in ps:
                it = self._item(p, ("pk", key))
                it["dst_t"], it["ld_t"] = buf[:, c:c + p.shape[0]], tot
                c += p.shape[0]
                self._fill_now({**it, "dst": None, "ld": 0})
            self._dirty = True
        return buf

    def clear(self):
        """Weights changed (optimizer step / state-dict load): every registered copy is re-cast on next use."""
        self._stale = True


# ------------------------------------------------------------------------------------------------ ViT
@dataclass
class _Blk:
    x: Tensor = None          # block input, fp32 [M,D]
    mean1: Tensor = None
    rstd1: Tensor = None
    qkv: Tensor = None        # bf16 [M,3D]
    o: Tensor = None          # bf16 [M,D]  attention output
    lse: Tensor = None        # fp32 [B,H,S]
    x1: Tensor = None         # fp32 [M,D]  after attention residual
    mean2: Tensor = None
    rstd2: Tensor = None
    h: Tensor = None          # bf16 [M,4D] fc1 pre-activation
    y: Tensor = None          # bf16 [M,D]  LN1 output (wgrad operand; 0.6 GB/layer at b=4096 is cheaper than re-reading x)
    y2: Tensor = None         # bf16 [M,D]  LN2 output
    a: Tensor = None          # bf16 [M,4D] gelu(h) when kept (keep_gelu_output), else re-emitted in backward


@dataclass
class VitSaved:
    B: int = 0
    S: int = 0
    patches: Tensor = None
    blocks: List[_Blk] = field(default_factory=list)
    x_last: Tensor = None
    mean_n: Tensor = None
    rstd_n: Tensor = None


def vit_forward(m, image: Tensor, wc: Bf16Weights, save: bool):
    """``ViTModel.forward`` (``simseg/models/backbones/mml/vit_builder.py:13-21``): all tokens, fp32 + bf16."""
    B = image.shape[0]
    D, H = m.embed_dim, m.num_heads
    N = m.patch_embed.num_patches
    S = N + 1
    M = B * S
    sv = VitSaved(B=B, S=S) if save else None
    patches = ops.im2col16(image.contiguous().float())
    pe = ops.linear_fwd(patches, wc.get(m.patch_embed.proj.weight), m.patch_embed.proj.bias)
    x = ops.vit_tokens_fwd(pe, m.cls_token.reshape(-1), m.pos_embed.reshape(S, D), B, N, D).reshape(M, D)
    if save:
        sv.patches = patches
    strides = (S * 3 * D, 3 * D, 64)
    # Residual adds are fused into the LayerNorm that follows them (ops.add_layernorm_fwd): every Linear writes its
    # bf16 output (+bias) through the coalesced TMA-store epilogue, exactly the tensor the reference's autocast produces,
    # and the fp32 residual stream is updated where it is read anyway.
    f = None                                      # bf16 output of the previous block's fc2, not yet added to x
    keep_a = save and keep_gelu_output(len(m.blocks) * M * 4 * D * 2, x.device)
    for blk in m.blocks:
        if f is None:
            y, _, mean1, rstd1 = ops.layernorm_fwd(x, blk.norm1.weight, blk.norm1.bias, 1e-6)
        else:
            x, y, _, mean1, rstd1 = ops.add_layernorm_fwd(x, f, blk.norm1.weight, blk.norm1.bias, 1e-6)
        qkv = ops.linear_fwd(y, wc.get(blk.attn.qkv.weight), blk.attn.qkv.bias)
        q5 = qkv.view(B, S, 3, H, 64)
        o, lse = ops.attention_fwd(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], B, H, S, strides, None, 0.125)
        o = o.view(M, D)
        pr = ops.linear_fwd(o, wc.get(blk.attn.proj.weight), blk.attn.proj.bias)
        x1, y2, _, mean2, rstd2 = ops.add_layernorm_fwd(x, pr, blk.norm2.weight, blk.norm2.bias, 1e-6)
        del pr
        h = torch.empty((M, 4 * D), device=x.device, dtype=torch.bfloat16) if save else None
        a = ops.linear_fwd(y2, wc.get(blk.mlp.fc1.weight), blk.mlp.fc1.bias, epilogue=EPI_BIAS_GELU, aux=h)
        f = ops.linear_fwd(a, wc.get(blk.mlp.fc2.weight), blk.mlp.fc2.bias)
        if save:
            sv.blocks.append(_Blk(x, mean1, rstd1, qkv, o, lse, x1, mean2, rstd2, h, y, y2, a if keep_a else None))
        del a
        x = x1
    if f is None:
        tok_bf16, tok_f32, mean_n, rstd_n = ops.layernorm_fwd(x, m.norm.weight, m.norm.bias, 1e-6, want_f32=True)
    else:
        x, tok_bf16, tok_f32, mean_n, rstd_n = ops.add_layernorm_fwd(x, f, m.norm.weight, m.norm.bias, 1e-6, want_f32=True)
    if save:
        sv.x_last, sv.mean_n, sv.rstd_n = x, mean_n, rstd_n
    return tok_f32.view(B, S, D), tok_bf16.view(B, S, D), sv


def vit_backward(m, sv: VitSaved, dtok: Tensor, wc: Bf16Weights, dtok2: Optional[Tensor] = None, grad_of=None):
    """Backward of ``vit_forward``.  ``dtok`` [B,S,D] bf16|f32 (+ optional fp32 ``dtok2``) = dL/d tokens.
    ``grad_of(param)`` names the fp32 buffer a parameter's gradient is ACCUMULATED into (default: ``param.grad``)."""
    _grad_of = grad_of or param_grad
    B, S = sv.B, sv.S
    D, H = m.embed_dim, m.num_heads
    M = B * S
    dev = dtok.device
    strides = (S * 3 * D, 3 * D, 64)
    dx = torch.empty((M, D), device=dev, dtype=torch.float32)
    g = torch.empty((M, D), device=dev, dtype=torch.bfloat16)
    nb = len(m.blocks)
    last_fc2_bias = _grad_of(m.blocks[-1].mlp.fc2.bias) if nb else None
    ops.layernorm_bwd(dtok.reshape(M, D), sv.x_last, m.norm.weight, sv.mean_n, sv.rstd_n,
                      dy2=None if dtok2 is None else dtok2.reshape(M, D), dx=dx, dx_bf16=g,
                      dgamma=_grad_of(m.norm.weight), dbeta=_grad_of(m.norm.bias), dx_colsum=last_fc2_bias)
    sv.x_last = None
    for i in range(nb - 1, -1, -1):
        blk, s = m.blocks[i], sv.blocks[i]
        # ---- MLP branch: x2 = x1 + fc2(gelu(fc1(LN2(x1))))
        # dgrad first: its epilogue multiplies by gelu'(h) AND emits a = gelu(h), which the fc2 wgrad then consumes
        a = s.a if s.a is not None else torch.empty_like(s.h)
        dh = ops.linear_dgrad(g, wc.get_t(blk.mlp.fc2.weight), epilogue=EPI_DGELU, aux=s.h, aux2=None if s.a is not None else a,
                              col_sum=_grad_of(blk.mlp.fc1.bias))
        s.h = s.a = None
        ops.linear_wgrad(g, a, _grad_of(blk.mlp.fc2.weight), accumulate=True)
        del a
        ops.linear_wgrad(dh, s.y2, _grad_of(blk.mlp.fc1.weight), accumulate=True)
        s.y2 = None
        dy2 = ops.linear_dgrad(dh, wc.get_t(blk.mlp.fc1.weight))
        del dh
        ops.layernorm_bwd(dy2, s.x1, blk.norm2.weight, s.mean2, s.rstd2, dx=dx, dx_accumulate=True, dx_bf16=g,
                          dgamma=_grad_of(blk.norm2.weight), dbeta=_grad_of(blk.norm2.bias),
                          dx_colsum=_grad_of(blk.attn.proj.bias))
        del dy2
        s.x1 = None
        # ---- attention branch: x1 = x + proj(attn(qkv(LN1(x))))
        ops.linear_wgrad(g, s.o, _grad_of(blk.attn.proj.weight), accumulate=True)
        do = ops.linear_dgrad(g, wc.get_t(blk.attn.proj.weight))
        dqkv = torch.empty_like(s.qkv)
        q5, d5 = s.qkv.view(B, S, 3, H, 64), dqkv.view(B, S, 3, H, 64)
        ops.attention_bwd(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], s.o, do, s.lse, B, H, S, strides, None, 0.125,
                          d5[:, :, 0], d5[:, :, 1], d5[:, :, 2])
        del do
        s.o = s.qkv = s.lse = None
        ops.colsum(dqkv, _grad_of(blk.attn.qkv.bias), accumulate=True)
        ops.linear_wgrad(dqkv, s.y, _grad_of(blk.attn.qkv.weight), accumulate=True)
        s.y = None
        dy = ops.linear_dgrad(dqkv, wc.get_t(blk.attn.qkv.weight))
        del dqkv
        prev_bias = _grad_of(m.blocks[i - 1].mlp.fc2.bias) if i > 0 else None
        ops.layernorm_bwd(dy, s.x, blk.norm1.weight, s.mean1, s.rstd1, dx=dx, dx_accumulate=True, dx_bf16=g,
                          dgamma=_grad_of(blk.norm1.weight), dbeta=_grad_of(blk.norm1.bias), dx_colsum=prev_bias)
        del dy
        s.x = None
    N = S - 1
    dpatch = ops.vit_tokens_bwd(dx, B, N, D, _grad_of(m.pos_embed).view(S, D), _grad_of(m.cls_token).view(D))
    ops.linear_wgrad(dpatch, sv.patches, _grad_of(m.patch_embed.proj.weight).view(D, 768), accumulate=True)
    ops.colsum(dpatch, _grad_of(m.patch_embed.proj.bias), accumulate=True)


# ------------------------------------------------------------------------------------------------ BERT
@dataclass
class _Lyr:
    h_in: Tensor = None       # bf16 [M,D] layer input (GEMM operand)
    qkv: Tensor = None        # bf16 [M,3D]
    c: Tensor = None          # bf16 [M,D]
    lse: Tensor = None
    s1: Tensor = None         # fp32 [M,D] attention.output pre-LN sum
    mean1: Tensor = None
    rstd1: Tensor = None
    pre: Tensor = None        # bf16 [M,F] intermediate pre-activation
    h1b: Tensor = None        # bf16 [M,D] attention.output LayerNorm result (wgrad operand)
    s2: Tensor = None         # fp32 [M,D] output pre-LN sum
    mean2: Tensor = None
    rstd2: Tensor = None
    act: Tensor = None        # bf16 [M,F] gelu(pre) when kept (keep_gelu_output), else re-emitted in backward


@dataclass
class BertDropout:
    """Train-mode dropout of HF ``BertModel`` (``hidden_dropout_prob`` / ``attention_probs_dropout_prob``, 0.1 each in
    bert-base-uncased; active under ``model.train()`` in the reference, ``huggingface_builder.py:16-17``).  ``rng`` is the
    DEVICE int64 pair {seed, step} the kernels read; a forward works on its own copy, which backward reuses, so the masks of
    the two passes agree whatever happens to the model's counter in between.  Sites: 0 = embeddings, then per layer l:
    1 + 3l attention probabilities, 2 + 3l attention.output dense, 3 + 3l output dense (the order HF calls them in)."""
    p_hidden: float = 0.0
    p_attn: float = 0.0
    rng: Tensor = None

    def hidden(self, site: int):
        return ops.Drop(self.p_hidden, self.rng, site) if self.p_hidden > 0 else None

    def attn(self, site: int):
        return ops.Drop(self.p_attn, self.rng, site) if self.p_attn > 0 else None


@dataclass
class BertSaved:
    B: int = 0
    T: int = 0
    ids: Tensor = None
    key_len: Tensor = None
    drop: Optional[BertDropout] = None
    e: Tensor = None
    mean0: Tensor = None
    rstd0: Tensor = None
    layers: List[_Lyr] = field(default_factory=list)


def _qkv_params(layer):
    a = layer.attention.self
    return [a.query, a.key, a.value]


def qkv_adjacent_order(named_params):
    """Parameters of the BERT tower in an order that puts each layer's query / key / value weights next to each other (and
    their biases likewise).  ``dist.FlatGrads`` lays gradients out in the order it is given, so with this order the three
    weight gradients of a layer are ONE contiguous [3D, D] matrix and ``bert_backward`` fills it with a single wgrad GEMM
    (and one column-sum launch for the biases) instead of three of each."""
    named = list(named_params)
    by_name = dict(named)
    out, done = [], set()
    for name, p in named:
        if name in done:
            continue
        if name.endswith("attention.self.query.weight"):
            base = name[:-len("query.weight")]
            for suffix in ("query.weight", "key.weight", "value.weight", "query.bias", "key.bias", "value.bias"):
                if base + suffix in by_name:
                    out.append(by_name[base + suffix])
                    done.add(base + suffix)
            continue
        out.append(p)
        done.add(name)
    return out


def _adjacent(ts: List[Tensor]) -> Optional[Tensor]:
    """One tensor over ``ts`` stacked along dim 0 if they sit back to back in one fp32 buffer (see ``qkv_adjacent_order``)."""
    t0 = ts[0]
    end = t0.data_ptr() + t0.numel() * 4
    for t in ts[1:]:
        if (t.dtype != torch.float32 or not t.is_contiguous() or t.shape[1:] != t0.shape[1:] or t.data_ptr() != end
                or t.untyped_storage().data_ptr() != t0.untyped_storage().data_ptr()):
            return None
        end += t.numel() * 4
    if t0.dtype != torch.float32 or not t0.is_contiguous():
        return None
    rows = sum(t.shape[0] for t in ts)
    size = (rows,) + tuple(t0.shape[1:])
    return torch.as_strided(t0, size, t0.stride())


def bert_forward(m, input_ids: Tensor, attention_mask: Tensor, wc: Bf16Weights, save: bool,
                 drop: Optional[BertDropout] = None):
    """HF ``BertModel(...).last_hidden_state`` as called by ``huggingface_builder.py:16-17``; ``drop`` = train-mode dropout
    (None in eval mode / p = 0): after the embedding LayerNorm, on the attention probabilities (keep bits drawn once per
    layer by ``ops.attn_dropout_mask``, applied inside the attention kernel after the softmax), and on the two dense
    outputs inside the residual-add LayerNorm kernels.

    The additive ``finfo.min`` key mask of a left-aligned ``attention_mask`` is applied as a per-sample key
    length inside the attention kernel."""
    B, T = input_ids.shape
    emb = m.embeddings
    D = emb.word_embeddings.weight.shape[1]
    H = m.num_heads
    M = B * T
    key_len = attention_mask.sum(1).to(torch.int32).contiguous()
    if drop is not None and drop.p_hidden <= 0 and drop.p_attn <= 0:
        drop = None
    sv = BertSaved(B=B, T=T, ids=input_ids.contiguous(), key_len=key_len, drop=drop) if save else None
    hid = (lambda site: drop.hidden(site)) if drop is not None else (lambda site: None)
    amask = None
    e = ops.bert_embed_fwd(input_ids.contiguous(), emb.word_embeddings.weight, emb.position_embeddings.weight,
                           emb.token_type_embeddings.weight).view(M, D)
    hb, hf, mean0, rstd0 = ops.layernorm_fwd(e, emb.LayerNorm.weight, emb.LayerNorm.bias, 1e-12, want_f32=True, drop=hid(0))
    if save:
        sv.e, sv.mean0, sv.rstd0 = e, mean0, rstd0
    strides = (T * 3 * D, 3 * D, 64)
    keep_a = save and keep_gelu_output(len(m.encoder.layer) * M * m.encoder.layer[0].intermediate.dense.weight.shape[0] * 2, e.device)
    for li, layer in enumerate(m.encoder.layer):
        qp = _qkv_params(layer)
        wqkv = wc.packed(f"bert.qkv.{li}", [p.weight for p in qp])
        bqkv = torch.cat([p.bias.detach() for p in qp])
        qkv = ops.linear_fwd(hb, wqkv, bqkv)
        q5 = qkv.view(B, T, 3, H, 64)
        if drop is not None and drop.p_attn > 0:
            amask = ops.attn_dropout_mask(B, H, T, drop.attn(1 + 3 * li), out=amask)      # one buffer, redrawn per layer
            c, lse = ops.attention_fwd(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], B, H, T, strides, key_len, 0.125,
                                       drop_mask=amask, drop_p=drop.p_attn)
        else:
            c, lse = ops.attention_fwd(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], B, H, T, strides, key_len, 0.125)
        c = c.view(M, D)
        ao = layer.attention.output
        d1 = ops.linear_fwd(c, wc.get(ao.dense.weight), ao.dense.bias)
        s1, h1b, h1f, mean1, rstd1 = ops.add_layernorm_fwd(hf, d1, ao.LayerNorm.weight, ao.LayerNorm.bias, 1e-12, want_f32=True,
                                                           drop=hid(2 + 3 * li))
        del d1
        F = layer.intermediate.dense.weight.shape[0]
        pre = torch.empty((M, F), device=e.device, dtype=torch.bfloat16) if save else None
        f = ops.linear_fwd(h1b, wc.get(layer.intermediate.dense.weight), layer.intermediate.dense.bias,
                           epilogue=EPI_BIAS_GELU, aux=pre)
        d2 = ops.linear_fwd(f, wc.get(layer.output.dense.weight), layer.output.dense.bias)
        act = f if keep_a else None
        del f
        s2, h2b, h2f, mean2, rstd2 = ops.add_layernorm_fwd(h1f, d2, layer.output.LayerNorm.weight, layer.output.LayerNorm.bias,
                                                           1e-12, want_f32=True, drop=hid(3 + 3 * li))
        del d2
        if save:
            sv.layers.append(_Lyr(hb, qkv, c, lse, s1, mean1, rstd1, pre, h1b, s2, mean2, rstd2, act))
        hb, hf = h2b, h2f
    return hf.view(B, T, D), hb.view(B, T, D), sv


def bert_backward(m, sv: BertSaved, dh: Tensor, wc: Bf16Weights, dh2: Optional[Tensor] = None, grad_of=None):
    """Backward of ``bert_forward``; ``dh`` [B,T,D] bf16|f32 (+ optional fp32 ``dh2``) = dL/d last_hidden_state.
    ``grad_of`` as in ``vit_backward``."""
    _grad_of = grad_of or param_grad
    B, T = sv.B, sv.T
    emb = m.embeddings
    D = emb.word_embeddings.weight.shape[1]
    H = m.num_heads
    M = B * T
    dev = dh.device
    strides = (T * 3 * D, 3 * D, 64)
    dy = dh.reshape(M, D)
    dres = None if dh2 is None else dh2.reshape(M, D)     # fp32 residual-path gradient added to dy
    drop = sv.drop                                        # the forward's {seed, step}: every mask is regenerated from it
    hid = (lambda site: drop.hidden(site)) if drop is not None else (lambda site: None)
    amask = None
    for li in range(len(m.encoder.layer) - 1, -1, -1):
        layer, s = m.encoder.layer[li], sv.layers[li]
        ao = layer.attention.output
        # h2 = LN(s2), s2 = out.dense(gelu(inter.dense(h1))) + h1
        ds2 = torch.empty((M, D), device=dev, dtype=torch.float32)
        g2 = torch.empty((M, D), device=dev, dtype=torch.bfloat16)
        ops.layernorm_bwd(dy, s.s2, layer.output.LayerNorm.weight, s.mean2, s.rstd2, dy2=dres, dx=ds2, dx_bf16=g2,
                          dgamma=_grad_of(layer.output.LayerNorm.weight), dbeta=_grad_of(layer.output.LayerNorm.bias),
                          dx_colsum=_grad_of(layer.output.dense.bias), drop=hid(3 + 3 * li), drop_mode=1)
        f = s.act if s.act is not None else torch.empty_like(s.pre)
        dpre = ops.linear_dgrad(g2, wc.get_t(layer.output.dense.weight), epilogue=EPI_DGELU, aux=s.pre,
                                aux2=None if s.act is not None else f, col_sum=_grad_of(layer.intermediate.dense.bias))
        s.act = None
        ops.linear_wgrad(g2, f, _grad_of(layer.output.dense.weight), accumulate=True)
        del f
        ops.linear_wgrad(dpre, s.h1b, _grad_of(layer.intermediate.dense.weight), accumulate=True)
        dh1 = ops.linear_dgrad(dpre, wc.get_t(layer.intermediate.dense.weight))
        del dpre
        # h1 = LN(s1), s1 = attn.out.dense(c) + h_in ; dh1_total = dh1 + ds2
        ds1 = torch.empty((M, D), device=dev, dtype=torch.float32)
        g1 = g2
        ops.layernorm_bwd(dh1, s.s1, ao.LayerNorm.weight, s.mean1, s.rstd1, dy2=ds2, dx=ds1, dx_bf16=g1,
                          dgamma=_grad_of(ao.LayerNorm.weight), dbeta=_grad_of(ao.LayerNorm.bias),
                          dx_colsum=_grad_of(ao.dense.bias), drop=hid(2 + 3 * li), drop_mode=1)
        del dh1, ds2
        ops.linear_wgrad(g1, s.c, _grad_of(ao.dense.weight), accumulate=True)
        dc = ops.linear_dgrad(g1, wc.get_t(ao.dense.weight))
        dqkv = torch.empty_like(s.qkv)
        q5, d5 = s.qkv.view(B, T, 3, H, 64), dqkv.view(B, T, 3, H, 64)
        if drop is not None and drop.p_attn > 0:
            amask = ops.attn_dropout_mask(B, H, T, drop.attn(1 + 3 * li), out=amask)
            ops.attention_bwd(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], s.c, dc, s.lse, B, H, T, strides, sv.key_len, 0.125,
                              d5[:, :, 0], d5[:, :, 1], d5[:, :, 2], drop_mask=amask, drop_p=drop.p_attn)
        else:
            ops.attention_bwd(q5[:, :, 0], q5[:, :, 1], q5[:, :, 2], s.c, dc, s.lse, B, H, T, strides, sv.key_len, 0.125,
                              d5[:, :, 0], d5[:, :, 1], d5[:, :, 2])
        del dc
        qp = _qkv_params(layer)
        gw, gb = [_grad_of(p.weight) for p in qp], [_grad_of(p.bias) for p in qp]
        gw_all, gb_all = _adjacent(gw), _adjacent(gb)
        if gb_all is not None:
            ops.colsum(dqkv, gb_all, accumulate=True)
        else:
            bsum = ops.colsum(dqkv)
            for j in range(3):
                gb[j].add_(bsum[j * D:(j + 1) * D])
        if gw_all is not None:                            # the three weight gradients are one [3D, D] matrix: one GEMM
            ops.linear_wgrad(dqkv, s.h_in, gw_all, accumulate=True)
        else:
            for j in range(3):
                ops.linear_wgrad(dqkv[:, j * D:(j + 1) * D], s.h_in, gw[j], accumulate=True)
        dy = ops.linear_dgrad(dqkv, wc.packed_t(f"bert.qkv.{li}", [p.weight for p in qp]))
        dres = ds1
        sv.layers[li] = None
    de = torch.empty((M, D), device=dev, dtype=torch.float32)
    ops.layernorm_bwd(dy, sv.e, emb.LayerNorm.weight, sv.mean0, sv.rstd0, dy2=dres, dx=de,
                      dgamma=_grad_of(emb.LayerNorm.weight), dbeta=_grad_of(emb.LayerNorm.bias), drop=hid(0), drop_mode=2)
    ops.bert_embed_bwd(sv.ids, de, _grad_of(emb.word_embeddings.weight), _grad_of(emb.position_embeddings.weight),
                       _grad_of(emb.token_type_embeddings.weight)[0])
