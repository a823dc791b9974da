"""Tensor-level wrappers over the C ABI.  PyTorch is used for device memory and streams only: every
wrapper passes raw device pointers + sizes + the current CUDA stream to ``libsimseg_b200.so``.
All functions raise on non-CUDA tensors — there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib
from ._lib import (BF16, EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_DGELU, EPI_NONE, EPI_ROWSCALE, F32,
                   PREC_FP32, PREC_SPLIT_BF16, PREC_TF32, GemmArgs, check)

Tensor = torch.Tensor
_ctx = {}


def ctx(device: Optional[int] = None) -> C.c_void_p:
    """Per-device simseg_ctx (created on first use)."""
    if not torch.cuda.is_available():
        raise _lib.SimsegError("simseg_b200 needs a CUDA device (B200, sm_100a); no CPU fallback exists")
    dev = torch.cuda.current_device() if device is None else device
    if dev not in _ctx:
        p = C.c_void_p()
        check(_lib.load().simseg_ctx_create(dev, C.byref(p)), "simseg_ctx_create")
        _ctx[dev] = p
    return _ctx[dev]


def launch_count(reset: bool = False) -> int:
    return int(_lib.load().simseg_ctx_launch_count(ctx(), 1 if reset else 0))


def _stream() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t: Optional[Tensor]) -> C.c_void_p:
    if t is None:
        return C.c_void_p(0)
    if not t.is_cuda:
        raise _lib.SimsegError("tensor is not on a CUDA device")
    return C.c_void_p(t.data_ptr())


def _dt(t: Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise _lib.SimsegError(f"unsupported dtype {t.dtype}")


def _rowmajor2d(t: Tensor):
    assert t.dim() == 2 and t.stride(1) == 1, f"need row-major 2-D tensor, got {tuple(t.shape)} {t.stride()}"
    return t.stride(0)


# --------------------------------------------------------------------------------------- GEMM
def gemm(a: Tensor, b: Tensor, *, M: int, N: int, K: int, a_major: int = 0, b_major: int = 0,
         out: Optional[Tensor] = None, out_dtype: torch.dtype = torch.bfloat16, epilogue: int = EPI_NONE,
         bias: Optional[Tensor] = None, residual: Optional[Tensor] = None, aux: Optional[Tensor] = None,
         row_scale: Optional[Tensor] = None, col_sum: Optional[Tensor] = None, accumulate: bool = False,
         aux2: Optional[Tensor] = None, tile_n: int = 0, _dbg: int = 0) -> Tensor:
    """D[M,N] = epilogue(sum_k A(m,k) B(n,k)).  a_major/b_major: 0 = operand stored [MN,K], 1 = stored [K,MN]."""
    lda, ldb = _rowmajor2d(a), _rowmajor2d(b)
    assert a.dtype == b.dtype
    if out is None:
        out = torch.empty((M, N), device=a.device, dtype=out_dtype)
    g = GemmArgs()
    g.a, g.b, g.d = a.data_ptr(), b.data_ptr(), out.data_ptr()
    g.M, g.N, g.K = M, N, K
    g.lda, g.ldb, g.ldd = lda, ldb, _rowmajor2d(out)
    g.a_major, g.b_major = a_major, b_major
    g.in_dtype, g.out_dtype = _dt(a), _dt(out)
    g.epilogue, g.accumulate = epilogue, 1 if accumulate else 0
    if bias is not None:
        assert bias.dtype == torch.float32 and bias.numel() == N
        g.bias = bias.data_ptr()
    if residual is not None:
        g.residual, g.ld_res, g.res_dtype = residual.data_ptr(), _rowmajor2d(residual), _dt(residual)
    if aux is not None:
        assert aux.dtype == torch.bfloat16
        g.aux, g.ld_aux = aux.data_ptr(), _rowmajor2d(aux)
    if aux2 is not None:
        assert aux2.dtype == torch.bfloat16
        g.aux2, g.ld_aux2 = aux2.data_ptr(), _rowmajor2d(aux2)
    if row_scale is not None:
        g.row_scale = row_scale.data_ptr()
    if col_sum is not None:
        assert col_sum.dtype == torch.float32 and col_sum.numel() == N
        g.col_sum = col_sum.data_ptr()
    g.tile_n = tile_n
    g.reserved = _dbg
    check(_lib.load().simseg_gemm(ctx(), C.byref(g), _stream()), "simseg_gemm")
    return out


def linear_fwd(x: Tensor, w: Tensor, bias: Optional[Tensor] = None, **kw) -> Tensor:
    """y = x @ w^T (+bias): x [M,K] bf16, w [N,K] bf16."""
    return gemm(x, w, M=x.shape[0], N=w.shape[0], K=x.shape[1], bias=bias, **kw)


def linear_dgrad(dy: Tensor, w_t: Tensor, **kw) -> Tensor:
    """dx = dy @ w: dy [M,N_out] bf16, w_t = w^T [K_in,N_out] bf16 (the per-step transposed bf16 weight copy, so dgrad
    runs the same K-major TMA path as forward)."""
    return gemm(dy, w_t, M=dy.shape[0], N=w_t.shape[0], K=dy.shape[1], a_major=0, b_major=0, **kw)


def linear_wgrad(dy: Tensor, x: Tensor, out: Tensor, accumulate: bool = False) -> Tensor:
    """dW[N,K] (+)= dy^T @ x: both operands MN-major with the token axis as the reduction."""
    return gemm(dy, x, M=dy.shape[1], N=x.shape[1], K=dy.shape[0], a_major=1, b_major=1, out=out,
                accumulate=accumulate)


# --------------------------------------------------------------------------------------- elementwise
def cast_bf16(src: Tensor, transpose_too: bool = False):
    rows = src.shape[0] if src.dim() > 1 else 1
    cols = src.numel() // rows
    s = src.contiguous()
    dst = torch.empty_like(s, dtype=torch.bfloat16)
    dst_t = torch.empty((cols, rows), device=s.device, dtype=torch.bfloat16) if transpose_too else None
    check(_lib.load().simseg_cast_bf16(ctx(), _p(s), _p(dst), _p(dst_t), rows, cols, _stream()), "cast_bf16")
    return (dst, dst_t) if transpose_too else dst


def colsum(x: Tensor, out: Optional[Tensor] = None, accumulate: bool = False) -> Tensor:
    M, N = x.shape
    if out is None:
        out = torch.empty(N, device=x.device, dtype=torch.float32)
        accumulate = False
    check(_lib.load().simseg_colsum(ctx(), _p(x), _dt(x), M, N, _rowmajor2d(x), _p(out), 1 if accumulate else 0,
                                    _stream()), "colsum")
    return out


def gelu_fwd(h: Tensor, out: Optional[Tensor] = None) -> Tensor:
    assert h.dtype == torch.bfloat16 and h.is_contiguous()
    out = torch.empty_like(h) if out is None else out
    check(_lib.load().simseg_gelu_fwd(ctx(), _p(h), _p(out), h.numel(), _stream()), "gelu_fwd")
    return out


class Drop:
    """One train-mode dropout site (HF BertModel: hidden / attention-probability dropout): probability ``p``, the DEVICE
    int64 pair ``rng`` = {seed, step} the kernels read, and the ``site`` number that separates the dropout layers of a step."""
    __slots__ = ("p", "rng", "site")

    def __init__(self, p: float, rng: Tensor, site: int):
        assert 0.0 < p < 1.0 and rng.dtype == torch.int64 and rng.numel() == 2 and rng.is_cuda
        self.p, self.rng, self.site = float(p), rng, int(site)


def layernorm_fwd(x: Tensor, gamma: Tensor, beta: Tensor, eps: float, *, want_bf16=True, want_f32=False,
                  want_stats=True, drop: Optional[Drop] = None):
    """``drop``: the OUTPUTS are dropped (HF BertEmbeddings: dropout(LayerNorm(e)))."""
    M, D = x.shape
    assert x.is_contiguous()
    yb = torch.empty((M, D), device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    yf = torch.empty((M, D), device=x.device, dtype=torch.float32) if want_f32 else None
    mean = torch.empty(M, device=x.device, dtype=torch.float32) if want_stats else None
    rstd = torch.empty(M, device=x.device, dtype=torch.float32) if want_stats else None
    if drop is not None:
        assert x.dtype == torch.float32
        check(_lib.load().simseg_layernorm_fwd_dropout(ctx(), _p(x), None, _p(gamma), _p(beta), eps, M, D, None, _p(yb), _p(yf),
                                                       _p(mean), _p(rstd), drop.p, _p(drop.rng), drop.site, _stream()),
              "layernorm_fwd_dropout")
        return yb, yf, mean, rstd
    check(_lib.load().simseg_layernorm_fwd(ctx(), _p(x), _dt(x), _p(gamma), _p(beta), eps, M, D, _p(yb), _p(yf),
                                           _p(mean), _p(rstd), _stream()), "layernorm_fwd")
    return yb, yf, mean, rstd


def add_layernorm_fwd(x: Tensor, add: Tensor, gamma: Tensor, beta: Tensor, eps: float, *, want_sum=True, want_bf16=True,
                      want_f32=False, want_stats=True, drop: Optional[Drop] = None):
    """s = x (f32) + add (bf16); returns (s, LN(s) bf16, LN(s) f32, mean, rstd) — residual add fused into the LayerNorm.
    ``drop``: s = x + dropout(add) (HF BertSelfOutput / BertOutput)."""
    M, D = x.shape
    assert x.is_contiguous() and add.is_contiguous() and x.dtype == torch.float32 and add.dtype == torch.bfloat16
    sm = torch.empty((M, D), device=x.device, dtype=torch.float32) if want_sum else None
    yb = torch.empty((M, D), device=x.device, dtype=torch.bfloat16) if want_bf16 else None
    yf = torch.empty((M, D), device=x.device, dtype=torch.float32) if want_f32 else None
    mean = torch.empty(M, device=x.device, dtype=torch.float32) if want_stats else None
    rstd = torch.empty(M, device=x.device, dtype=torch.float32) if want_stats else None
    if drop is not None:
        check(_lib.load().simseg_layernorm_fwd_dropout(ctx(), _p(x), _p(add), _p(gamma), _p(beta), eps, M, D, _p(sm), _p(yb),
                                                       _p(yf), _p(mean), _p(rstd), drop.p, _p(drop.rng), drop.site, _stream()),
              "layernorm_fwd_dropout")
        return sm, yb, yf, mean, rstd
    check(_lib.load().simseg_add_layernorm_fwd(ctx(), _p(x), _p(add), _p(gamma), _p(beta), eps, M, D, _p(sm), _p(yb), _p(yf),
                                               _p(mean), _p(rstd), _stream()), "add_layernorm_fwd")
    return sm, yb, yf, mean, rstd


def layernorm_bwd(dy: Tensor, x: Tensor, gamma: Tensor, mean: Tensor, rstd: Tensor, *, dy2: Optional[Tensor] = None,
                  dx: Optional[Tensor] = None, dx_accumulate: bool = False, dx_bf16: Optional[Tensor] = None,
                  dgamma: Optional[Tensor] = None, dbeta: Optional[Tensor] = None,
                  dx_colsum: Optional[Tensor] = None, drop: Optional[Drop] = None, drop_mode: int = 0):
    """``drop`` + ``drop_mode``: 1 = backward of ``add_layernorm_fwd(drop=)`` (dx_bf16 / dx_colsum carry the mask),
    2 = backward of ``layernorm_fwd(drop=)`` (dy + dy2 is masked first)."""
    M, D = x.shape
    assert dy.is_contiguous() and x.is_contiguous()
    if drop is not None:
        assert drop_mode in (1, 2) and x.dtype == torch.float32
        check(_lib.load().simseg_layernorm_bwd_dropout(ctx(), _p(dy), _dt(dy), _p(dy2), _p(x), _p(gamma), _p(mean), _p(rstd), M, D,
                                                       _p(dx), 1 if dx_accumulate else 0, _p(dx_bf16), _p(dgamma), _p(dbeta),
                                                       _p(dx_colsum), drop_mode, drop.p, _p(drop.rng), drop.site, _stream()),
              "layernorm_bwd_dropout")
        return
    check(_lib.load().simseg_layernorm_bwd(ctx(), _p(dy), _dt(dy), _p(dy2), _p(x), _dt(x), _p(gamma), _p(mean),
                                           _p(rstd), M, D, _p(dx), 1 if dx_accumulate else 0, _p(dx_bf16),
                                           _p(dgamma), _p(dbeta), _p(dx_colsum), _stream()), "layernorm_bwd")


# --------------------------------------------------------------------------------------- attention
def attention_fwd(q: Tensor, k: Tensor, v: Tensor, B: int, H: int, S: int, strides, key_len: Optional[Tensor],
                  scale: float, out: Optional[Tensor] = None, lse: Optional[Tensor] = None,
                  drop_mask: Optional[Tensor] = None, drop_p: float = 0.0):
    """``drop_mask`` (from ``attn_dropout_mask``) + ``drop_p``: out = dropout(softmax(..)) V (HF BertSelfAttention, train mode)."""
    sb, ss, sh = strides
    out = torch.empty((B, S, H * 64), device=q.device, dtype=torch.bfloat16) if out is None else out
    lse = torch.empty((B, H, S), device=q.device, dtype=torch.float32) if lse is None else lse
    if drop_mask is not None:
        check(_lib.load().simseg_attention_fwd_dropout(ctx(), _p(q), _p(k), _p(v), sb, ss, sh, B, H, S, _p(key_len), scale,
                                                       _p(out), _p(lse), _p(drop_mask), drop_p, _stream()), "attention_fwd_dropout")
        return out, lse
    check(_lib.load().simseg_attention_fwd(ctx(), _p(q), _p(k), _p(v), sb, ss, sh, B, H, S, _p(key_len), scale,
                                           _p(out), _p(lse), _stream()), "attention_fwd")
    return out, lse


def attn_dropout_mask(B: int, H: int, S: int, drop: Drop, out: Optional[Tensor] = None) -> Tensor:
    """Keep bits of one attention layer's probability dropout, in the attention kernels' tile coordinates (int32 words)."""
    n = _lib.load().simseg_attn_dropout_mask_words(B, H, S)
    if n <= 0:
        raise _lib.SimsegError(f"attention dropout: B={B} H={H} S={S} unsupported (S <= 224)")
    mask = torch.empty(n, device=drop.rng.device, dtype=torch.int32) if out is None else out
    assert mask.numel() >= n and mask.dtype == torch.int32
    check(_lib.load().simseg_attn_dropout_mask(ctx(), B, H, S, drop.p, _p(drop.rng), drop.site, _p(mask), mask.numel(), _stream()),
          "attn_dropout_mask")
    return mask


def attention_bwd(q, k, v, out, dout, lse, B, H, S, strides, key_len, scale, dq, dk, dv, drop_mask=None, drop_p=0.0):
    sb, ss, sh = strides
    if drop_mask is not None:
        check(_lib.load().simseg_attention_bwd_dropout(ctx(), _p(q), _p(k), _p(v), _p(out), _p(dout), _p(lse), sb, ss, sh, B, H, S,
                                                       _p(key_len), scale, _p(dq), _p(dk), _p(dv), _p(drop_mask), drop_p,
                                                       _stream()), "attention_bwd_dropout")
        return
    check(_lib.load().simseg_attention_bwd(ctx(), _p(q), _p(k), _p(v), _p(out), _p(dout), _p(lse), sb, ss, sh, B, H, S,
                                           _p(key_len), scale, _p(dq), _p(dk), _p(dv), _stream()), "attention_bwd")


# --------------------------------------------------------------------------------------- embeddings
def im2col16(image: Tensor) -> Tensor:
    B, Cc, Hi, Wi = image.shape
    assert Cc == 3 and image.dtype == torch.float32 and image.is_contiguous()
    out = torch.empty((B * (Hi // 16) * (Wi // 16), 768), device=image.device, dtype=torch.bfloat16)
    check(_lib.load().simseg_im2col16(ctx(), _p(image), B, Hi, Wi, _p(out), _stream()), "im2col16")
    return out


def vit_tokens_fwd(patch: Tensor, cls: Tensor, pos: Tensor, B: int, N: int, D: int) -> Tensor:
    x = torch.empty((B, N + 1, D), device=patch.device, dtype=torch.float32)
    check(_lib.load().simseg_vit_tokens_fwd(ctx(), _p(patch), _dt(patch), _p(cls), _p(pos), B, N, D, _p(x), _stream()),
          "vit_tokens_fwd")
    return x


def vit_tokens_bwd(dx: Tensor, B: int, N: int, D: int, dpos: Tensor, dcls: Tensor) -> Tensor:
    dpatch = torch.empty((B * N, D), device=dx.device, dtype=torch.bfloat16)
    check(_lib.load().simseg_vit_tokens_bwd(ctx(), _p(dx), B, N, D, _p(dpatch), _p(dpos), _p(dcls), _stream()),
          "vit_tokens_bwd")
    return dpatch


def bert_embed_fwd(ids: Tensor, word: Tensor, pos: Tensor, type_emb: Tensor) -> Tensor:
    B, T = ids.shape
    D = word.shape[1]
    assert ids.dtype == torch.int64 and ids.is_contiguous()
    e = torch.empty((B, T, D), device=ids.device, dtype=torch.float32)
    check(_lib.load().simseg_bert_embed_fwd(ctx(), _p(ids), _p(word), _p(pos), _p(type_emb), B, T, D, _p(e), _stream()),
          "bert_embed_fwd")
    return e


def bert_embed_bwd(ids: Tensor, de: Tensor, dword: Tensor, dpos: Tensor, dtype0: Tensor) -> None:
    B, T = ids.shape
    D = de.shape[-1]
    check(_lib.load().simseg_bert_embed_bwd(ctx(), _p(ids), _p(de), B, T, D, _p(dword), _p(dpos), _p(dtype0), _stream()),
          "bert_embed_bwd")


# --------------------------------------------------------------------------------------- heads
def topk_pool_l2norm_fwd(x: Tensor, k: int, tok_begin: int, ntok: int, attention_mask: Optional[Tensor] = None,
                         eps: float = 1e-8, l2norm: bool = True, save_idx: bool = True):
    B, S, E = x.shape
    assert x.is_contiguous()
    pooled = torch.empty((B, E), device=x.device, dtype=torch.float32)
    emb = torch.empty((B, E), device=x.device, dtype=torch.float32) if l2norm else None
    idx = torch.empty((B, k, E), device=x.device, dtype=torch.int32) if save_idx else None
    mask_ld = 0
    if attention_mask is not None:
        assert attention_mask.dtype == torch.int64 and attention_mask.stride(1) == 1
        mask_ld = attention_mask.stride(0)
    check(_lib.load().simseg_topk_pool_l2norm_fwd(ctx(), _p(x), _dt(x), B, S, E, tok_begin, ntok, k, _p(attention_mask),
                                                  mask_ld, eps, _p(pooled), _p(emb), _p(idx), _stream()),
          "topk_pool_l2norm_fwd")
    return pooled, emb, idx


def topk_pool_l2norm_bwd(demb: Tensor, pooled: Tensor, idx: Tensor, S: int, k: int, eps: float = 1e-8,
                         l2norm: bool = True, out: Optional[Tensor] = None) -> Tensor:
    B, E = demb.shape
    dx = torch.empty((B, S, E), device=demb.device, dtype=torch.bfloat16) if out is None else out
    check(_lib.load().simseg_topk_pool_l2norm_bwd(ctx(), _p(demb), _p(pooled), _p(idx), B, S, E, 0, k, eps,
                                                  1 if l2norm else 0, _p(dx), _stream()), "topk_pool_l2norm_bwd")
    return dx


def proj_topk_supported(S: int, D: int, E: int, k: int) -> bool:
    """Shapes the fused projection + top-k head (``simseg_proj_topk_*``) takes; others use the two-kernel path."""
    return S <= 256 and D % 128 == 0 and E % 128 == 0 and 1 <= k <= 8


def proj_topk_fwd(x: Tensor, w: Tensor, k: int, tok_begin: int, ntok: int, attention_mask: Optional[Tensor] = None,
                  eps: float = 1e-8, l2norm: bool = True, save_idx: bool = True):
    """x bf16 [B,S,D] (all tokens), w bf16 [E,D] -> (pooled, emb | None, idx | None) without the [B,S,E] projection."""
    B, S, D = x.shape
    E = w.shape[0]
    assert x.is_contiguous() and w.is_contiguous() and x.dtype == torch.bfloat16 and w.dtype == torch.bfloat16
    pooled = torch.empty((B, E), device=x.device, dtype=torch.float32)
    emb = torch.empty((B, E), device=x.device, dtype=torch.float32) if l2norm else None
    idx = torch.empty((B, k, E), device=x.device, dtype=torch.int32) if save_idx else None
    mask_ld = 0
    if attention_mask is not None:
        assert attention_mask.dtype == torch.int64 and attention_mask.stride(1) == 1
        mask_ld = attention_mask.stride(0)
    check(_lib.load().simseg_proj_topk_fwd(ctx(), _p(x), _p(w), B, S, D, E, tok_begin, ntok, k, _p(attention_mask), mask_ld,
                                           eps, _p(pooled), _p(emb), _p(idx), _stream()), "proj_topk_fwd")
    return pooled, emb, idx


def proj_topk_bwd(demb: Tensor, pooled: Tensor, idx: Tensor, x: Tensor, wt: Optional[Tensor], k: int, eps: float = 1e-8,
                  l2norm: bool = True, dw: Optional[Tensor] = None, want_dx: bool = True):
    """Backward of ``proj_topk_fwd``: returns dx f32 [B,S,D] (or None); ``dw`` f32 [E,D] is ACCUMULATED into when given.
    ``wt`` = bf16 [D,E] transposed weight copy (needed for dx)."""
    B, S, D = x.shape
    E = demb.shape[1]
    assert x.is_contiguous() and demb.is_contiguous() and demb.dtype == torch.float32
    gy = torch.empty((B, E), device=x.device, dtype=torch.float32)
    dx = torch.empty((B, S, D), device=x.device, dtype=torch.float32) if want_dx else None
    if dw is not None:
        assert dw.dtype == torch.float32 and dw.is_contiguous() and tuple(dw.shape) == (E, D)
    check(_lib.load().simseg_proj_topk_bwd(ctx(), _p(demb), _p(pooled), _p(idx), _p(x), _p(wt), B, S, D, E, k, eps,
                                           1 if l2norm else 0, _p(gy), _p(dx), _p(dw), _stream()), "proj_topk_bwd")
    return dx


# --------------------------------------------------------------------------------------- loss / similarity
def infonce_fwd(feat1: Tensor, feat2g: Tensor, temperature: Tensor, row_offset: int, precision: int = PREC_FP32,
                want_logits: bool = False):
    b, E = feat1.shape
    Bg = feat2g.shape[0]
    assert feat1.dtype == torch.float32 and feat2g.dtype == torch.float32 and feat1.is_contiguous() and feat2g.is_contiguous()
    dev = feat1.device
    cos = torch.empty((b, Bg), device=dev, dtype=torch.float32)
    logits = torch.empty((b, Bg), device=dev, dtype=torch.float32) if want_logits else None
    loss_rows = torch.empty(b, device=dev, dtype=torch.float32)
    lse = torch.empty(b, device=dev, dtype=torch.float32)
    argmax = torch.empty(b, device=dev, dtype=torch.int32)
    check(_lib.load().simseg_infonce_fwd(ctx(), _p(feat1), _p(feat2g), b, Bg, E, _p(temperature), row_offset, precision,
                                         _p(cos), _p(logits), _p(loss_rows), _p(lse), _p(argmax), _stream()),
          "infonce_fwd")
    return loss_rows, lse, argmax, cos, logits


def infonce_bwd(feat1: Tensor, feat2g: Tensor, temperature: Tensor, row_offset: int, lse: Tensor, grad_scale: float,
                cos: Tensor, dfeat2g: Tensor, dtemp: Tensor, precision: int = PREC_FP32) -> Tensor:
    b, E = feat1.shape
    Bg = feat2g.shape[0]
    dfeat1 = torch.empty_like(feat1)
    check(_lib.load().simseg_infonce_bwd(ctx(), _p(feat1), _p(feat2g), b, Bg, E, _p(temperature), row_offset, precision,
                                         _p(lse), grad_scale, _p(cos), _p(dfeat1), _p(dfeat2g), _p(dtemp), _stream()),
          "infonce_bwd")
    return dfeat1


def _workspace(nbytes: int, device) -> Tensor:
    """256-byte aligned caller-owned scratch (torch's caching allocator hands out 512-byte aligned blocks)."""
    ws = torch.empty(max(int(nbytes), 256), device=device, dtype=torch.uint8)
    assert ws.data_ptr() % 256 == 0
    return ws


def infonce_fused_fwd(feat1: Tensor, feat2g: Tensor, temperature: Tensor, row_offset: int):
    """NCE.forward on the tensor cores (split-bf16 products, scores consumed in the GEMM epilogue): returns
    (loss_rows [b], lse [b], argmax [b] int32, workspace) — hand the workspace to ``infonce_fused_bwd`` unchanged."""
    b, E = feat1.shape
    Bg = feat2g.shape[0]
    assert feat1.dtype == torch.float32 and feat2g.dtype == torch.float32 and feat1.is_contiguous() and feat2g.is_contiguous()
    dev = feat1.device
    lib = _lib.load()
    ws = _workspace(lib.simseg_infonce_fused_workspace_bytes(b, Bg, E), dev)
    loss_rows = torch.empty(b, device=dev, dtype=torch.float32)
    lse = torch.empty(b, device=dev, dtype=torch.float32)
    argmax = torch.empty(b, device=dev, dtype=torch.int32)
    check(lib.simseg_infonce_fused_fwd(ctx(), _p(feat1), _p(feat2g), b, Bg, E, _p(temperature), row_offset, _p(ws), ws.numel(),
                                       _p(loss_rows), _p(lse), _p(argmax), _stream()), "infonce_fused_fwd")
    return loss_rows, lse, argmax, ws


def infonce_fused_bwd(b: int, Bg: int, E: int, temperature: Tensor, row_offset: int, lse: Tensor, grad_scale: float,
                      ws: Tensor, dfeat2g: Optional[Tensor], dtemp: Optional[Tensor], want_dfeat1: bool = True):
    """dfeat1 [b,E] (returned), dfeat2g [Bg,E] accumulated, dtemp accumulated; consumes the forward's workspace."""
    dfeat1 = torch.empty((b, E), device=ws.device, dtype=torch.float32) if want_dfeat1 else None
    check(_lib.load().simseg_infonce_fused_bwd(ctx(), b, Bg, E, _p(temperature), row_offset, _p(lse), grad_scale, _p(ws),
                                               ws.numel(), _p(dfeat1), _p(dfeat2g), _p(dtemp), _stream()), "infonce_fused_bwd")
    return dfeat1


def retrieval_rank_fused(left: Tensor, right: Tensor, left_gid: Tensor, right_gid: Tensor) -> Tensor:
    """First-match rank of every left row (``EmbANN._ann`` + ``RetrievalMetric``, tasks/clip/hooks/utils.py:35-42,63-65) in one
    tensor-core pass; neither the [M,Nr] similarity matrix nor its argsort is materialised.  -1 where no right item matches."""
    M, E = left.shape
    Nr = right.shape[0]
    assert left.dtype == torch.float32 and right.dtype == torch.float32 and left.is_contiguous() and right.is_contiguous()
    assert left_gid.dtype == torch.int64 and right_gid.dtype == torch.int64 and left_gid.numel() == M and right_gid.numel() == Nr
    lib = _lib.load()
    ws = _workspace(lib.simseg_retrieval_fused_workspace_bytes(M, Nr, E), left.device)
    rank = torch.empty(M, device=left.device, dtype=torch.int32)
    check(lib.simseg_retrieval_rank_fused(ctx(), _p(left), _p(right), M, Nr, E, _p(left_gid.contiguous()), _p(right_gid.contiguous()),
                                          _p(ws), ws.numel(), _p(rank), _stream()), "retrieval_rank_fused")
    return rank


def allpairs_sim_split(left: Tensor, right: Tensor) -> Tensor:
    """left @ right^T (fp32 out) through the split-bf16 tensor-core products (|err| <= ~1e-5 on unit-norm rows)."""
    M, E = left.shape
    Nr = right.shape[0]
    assert left.dtype == torch.float32 and right.dtype == torch.float32 and left.is_contiguous() and right.is_contiguous()
    lib = _lib.load()
    ws = _workspace(lib.simseg_retrieval_fused_workspace_bytes(M, Nr, E), left.device)
    out = torch.empty((M, Nr), device=left.device, dtype=torch.float32)
    check(lib.simseg_allpairs_sim_split(ctx(), _p(left), _p(right), M, Nr, E, _p(ws), ws.numel(), _p(out), _stream()),
          "allpairs_sim_split")
    return out


def patch_text_sim(patches: Tensor, text: Tensor, normalize: bool = True, want_argmax: bool = True):
    """patches [..., E] (f32 | bf16), text [C,E] same dtype -> sim [..., C] f32, argmax [...] int32."""
    E = patches.shape[-1]
    lead = patches.shape[:-1]
    p2 = patches.reshape(-1, E)
    assert p2.is_contiguous() and text.is_contiguous() and text.dtype == patches.dtype
    rows, Cn = p2.shape[0], text.shape[0]
    sim = torch.empty((rows, Cn), device=p2.device, dtype=torch.float32)
    am = torch.empty(rows, device=p2.device, dtype=torch.int32) if want_argmax else None
    wsb = int(_lib.load().simseg_patch_text_sim_workspace_bytes(rows))
    ws = torch.empty(wsb, device=p2.device, dtype=torch.uint8)
    check(_lib.load().simseg_patch_text_sim(ctx(), _p(p2), _dt(p2), rows, E, _p(text), Cn, 1 if normalize else 0, _p(sim),
                                            _p(am), _p(ws), wsb, _stream()), "patch_text_sim")
    return sim.reshape(*lead, Cn), (am.reshape(*lead) if am is not None else None)


def allpairs_sim(left: Tensor, right: Tensor, precision: int = PREC_FP32) -> Tensor:
    M, E = left.shape
    Nr = right.shape[0]
    assert left.dtype == torch.float32 and right.dtype == torch.float32 and left.is_contiguous() and right.is_contiguous()
    out = torch.empty((M, Nr), device=left.device, dtype=torch.float32)
    check(_lib.load().simseg_allpairs_sim(ctx(), _p(left), _p(right), M, Nr, E, precision, _p(out), _stream()),
          "allpairs_sim")
    return out


def retrieval_rank(sim: Tensor, left_gid: Tensor, right_gid: Tensor) -> Tensor:
    M, Nr = sim.shape
    assert sim.is_contiguous() and left_gid.dtype == torch.int64 and right_gid.dtype == torch.int64
    rank = torch.empty(M, device=sim.device, dtype=torch.int32)
    check(_lib.load().simseg_retrieval_rank(ctx(), _p(sim), M, Nr, _p(left_gid), _p(right_gid), _p(rank), _stream()),
          "retrieval_rank")
    return rank


# --------------------------------------------------------------------------------------- zero-shot segmentation glue
def seg_class_embed(prompt_emb: Tensor) -> Tensor:
    """(C,P,E) fp32 prompt embeddings -> (C,E): mean over prompts, /= norm (``tools/seg_evaluation.py:71-72``)."""
    Cn, P, E = prompt_emb.shape
    x = prompt_emb.contiguous().float()
    out = torch.empty((Cn, E), device=x.device, dtype=torch.float32)
    check(_lib.load().simseg_seg_class_embed(ctx(), _p(x), Cn, P, E, _p(out), _stream()), "seg_class_embed")
    return out


def seg_select(img_emb: Tensor, text_emb: Tensor, top_cls_num: int, max_cand: int = 5):
    """Image-level class scores, top-k threshold and candidate classes (``tools/seg_evaluation.py:119-128,141-144``).
    Returns (scores [B,C], cand int32 [B,max_cand] padded with -1, threshold [B])."""
    B, E = img_emb.shape
    Cn = text_emb.shape[0]
    x, t = img_emb.contiguous().float(), text_emb.contiguous().float()
    scores = torch.empty((B, Cn), device=x.device, dtype=torch.float32)
    cand = torch.empty((B, max_cand), device=x.device, dtype=torch.int32)
    thr = torch.empty(B, device=x.device, dtype=torch.float32)
    check(_lib.load().simseg_seg_select(ctx(), _p(x), _p(t), B, Cn, E, top_cls_num, max_cand, _p(scores), _p(cand), _p(thr),
                                        _stream()), "seg_select")
    return scores, cand, thr


def seg_upsample_norm(sim: Tensor, cand: Tensor, h: int, w: int, scale: int = 16) -> Tensor:
    """sim [B,N,C] fp32 + candidates [B,K] -> min-max normalised, nearest x`scale` up-sampled maps [B,K,h*scale,w*scale]
    (``tools/seg_evaluation.py:136-139,146-147``)."""
    B, N, Cn = sim.shape
    K = cand.shape[1]
    s = sim.contiguous()
    assert s.dtype == torch.float32 and cand.dtype == torch.int32
    out = torch.empty((B, K, h * scale, w * scale), device=s.device, dtype=torch.float32)
    check(_lib.load().simseg_seg_upsample_norm(ctx(), _p(s), _p(cand.contiguous()), B, N, Cn, K, h, w, scale, _p(out), _stream()),
          "seg_upsample_norm")
    return out


def pos_embed_bicubic(pos_embed: Tensor, grid_dst: int, num_extra: int = 1) -> Tensor:
    """pos_embed [1, num_extra + g0*g0, D] fp32 -> [1, num_extra + grid_dst**2, D]: bicubic (align_corners=False) resize of the
    grid part, extra tokens copied (``simseg/utils/interpolate_pe.py:14-24``)."""
    assert pos_embed.dim() == 3 and pos_embed.shape[0] == 1 and pos_embed.dtype == torch.float32
    src = pos_embed.contiguous()
    D = src.shape[-1]
    g0 = int((src.shape[-2] - num_extra) ** 0.5)
    assert g0 * g0 + num_extra == src.shape[-2], "position embedding grid must be square"
    out = torch.empty((1, num_extra + grid_dst * grid_dst, D), device=src.device, dtype=torch.float32)
    check(_lib.load().simseg_pos_embed_bicubic(ctx(), _p(src), _p(out), g0, grid_dst, D, num_extra, _stream()), "pos_embed_bicubic")
    return out
