"""Per-op parity of the CUDA kernels (through the C ABI) against plain fp32 torch math on the same inputs.

Tolerances: bf16-input tensor-core ops are compared with an fp32 reference computed from the SAME bf16-rounded
inputs, so the only differences are accumulation order and the final bf16 rounding of the output
(<= 2^-8 relative); fp32 ops are held to 1e-5..1e-4.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


def _ops():
    from simseg_b200 import ops
    return ops


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12)).item()


def gelu(x):
    return 0.5 * x * (1 + torch.erf(x / math.sqrt(2)))


@pytest.mark.parametrize("M,N,K,tile_n", [(256, 128, 64, 128), (128, 256, 128, 256), (300, 384, 384, 0),
                                          (1000, 1152, 384, 192), (777, 512, 768, 0), (4096, 1536, 384, 0),
                                          (130, 171, 512, 192), (64, 20, 512, 0), (20000, 768, 200, 256)])
@pytest.mark.parametrize("mode", [16, 32], ids=["cta1", "pair"])
def test_gemm_kk_plain(cuda, M, N, K, tile_n, mode):
    """mode: 16 = single-CTA 128-row tiles, 32 = CTA pairs (tcgen05 cta_group::2, 256-row tiles)."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    a = torch.randn(M, K, device=cuda, generator=g).bfloat16()
    b = torch.randn(N, K, device=cuda, generator=g).bfloat16()
    bias = torch.randn(N, device=cuda, generator=g)
    ref = a.float() @ b.float().T + bias
    out = ops.gemm(a, b, M=M, N=N, K=K, bias=bias, out_dtype=torch.float32, tile_n=tile_n, _dbg=mode)
    torch.cuda.synchronize()
    assert _rel(out, ref) < 2e-5
    out16 = ops.gemm(a, b, M=M, N=N, K=K, bias=bias, out_dtype=torch.bfloat16, tile_n=tile_n, _dbg=mode)
    assert _rel(out16, ref) < 6e-3


@pytest.mark.parametrize("a_major,b_major", [(0, 1), (1, 1), (1, 0)])
@pytest.mark.parametrize("M,N,K", [(256, 128, 128), (333, 384, 200), (384, 1536, 3000), (512, 768, 64)])
@pytest.mark.parametrize("mode", [16, 32], ids=["cta1", "pair"])
def test_gemm_majors(cuda, a_major, b_major, M, N, K, mode):
    ops = _ops()
    K8 = (K + 7) // 8 * 8
    g = torch.Generator(device="cuda").manual_seed(7 * M + N + K)
    A = torch.randn(M, K8, device=cuda, generator=g).bfloat16()
    Bm = torch.randn(N, K8, device=cuda, generator=g).bfloat16()
    ref = A.float() @ Bm.float().T
    M8, N8 = (M + 7) // 8 * 8, (N + 7) // 8 * 8
    if a_major:                                   # store as [K, M] with 16-byte aligned rows
        a_st = torch.zeros(K8, M8, device=cuda, dtype=torch.bfloat16); a_st[:, :M] = A.T
    else:
        a_st = A
    if b_major:
        b_st = torch.zeros(K8, N8, device=cuda, dtype=torch.bfloat16); b_st[:, :N] = Bm.T
    else:
        b_st = Bm
    out = ops.gemm(a_st, b_st, M=M, N=N, K=K8, a_major=a_major, b_major=b_major, out_dtype=torch.float32, _dbg=mode)
    assert _rel(out, ref) < 2e-5


def test_gemm_splitk_accumulate(cuda):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(3)
    T, Do, Di = 20000, 384, 384                     # wgrad shape: K = tokens
    dy = torch.randn(T, Do, device=cuda, generator=g).bfloat16()
    x = torch.randn(T, Di, device=cuda, generator=g).bfloat16()
    ref = dy.float().T @ x.float()
    out = torch.empty(Do, Di, device=cuda)
    ops.linear_wgrad(dy, x, out)
    assert _rel(out, ref) < 1e-4
    ops.linear_wgrad(dy, x, out, accumulate=True)
    assert _rel(out, 2 * ref) < 1e-4


def test_gemm_epilogues(cuda):
    ops = _ops()
    from simseg_b200._lib import EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_DGELU, EPI_ROWSCALE
    g = torch.Generator(device="cuda").manual_seed(11)
    M, N, K = 1000, 1536, 384
    a = (torch.randn(M, K, device=cuda, generator=g) * 0.5).bfloat16()
    b = (torch.randn(N, K, device=cuda, generator=g) * 0.1).bfloat16()
    bias = torch.randn(N, device=cuda, generator=g) * 0.1
    pre = a.float() @ b.float().T + bias
    # GELU: pre-activation saved in bf16, activation computed from the rounded value
    aux = torch.empty(M, N, device=cuda, dtype=torch.bfloat16)
    act = ops.gemm(a, b, M=M, N=N, K=K, bias=bias, epilogue=EPI_BIAS_GELU, aux=aux)
    assert _rel(aux, pre) < 6e-3
    assert (act.float() - gelu(aux.float())).abs().max().item() < 2e-2
    assert _rel(act, gelu(pre)) < 1e-2
    # residual (fp32 stream)
    res = torch.randn(M, N, device=cuda, generator=g)
    out = ops.gemm(a, b, M=M, N=N, K=K, bias=bias, epilogue=EPI_BIAS_RESIDUAL, residual=res, out_dtype=torch.float32)
    assert _rel(out, pre + res) < 2e-5
    # dGELU + column sums
    cs = torch.zeros(N, device=cuda)
    dh = ops.gemm(a, b, M=M, N=N, K=K, epilogue=EPI_DGELU, aux=aux, col_sum=cs, out_dtype=torch.float32)
    h = aux.float().requires_grad_(True)
    gelu(h).backward(pre - bias)
    assert _rel(dh, h.grad) < 1e-4
    assert _rel(cs, h.grad.sum(0)) < 1e-4
    # row scale
    rs = torch.rand(M, device=cuda, generator=g) + 0.5
    o = ops.gemm(a, b, M=M, N=N, K=K, epilogue=EPI_ROWSCALE, row_scale=rs, out_dtype=torch.float32)
    assert _rel(o, (pre - bias) * rs[:, None]) < 2e-5


@pytest.mark.parametrize("M,N,K,tile_n", [(1000, 1536, 384, 0), (130, 200, 64, 128), (4500, 3072, 768, 256), (257, 1544, 128, 192),
                                          (40000, 1536, 384, 256)])
@pytest.mark.parametrize("mode", [16, 32], ids=["cta1", "pair"])
def test_gemm_tma_epilogue_bf16(cuda, M, N, K, tile_n, mode):
    """The coalesced (swizzled smem + TMA store) epilogue for bf16 outputs: plain+bias, GELU with saved pre-activation,
    and the backward dGELU that TMA-loads the pre-activation, emits gelu(pre) next to the gradient and column sums —
    ragged M / N (TMA clips), more tiles than SMs (staging-slot recycling across tiles), against the direct-store path."""
    ops = _ops()
    from simseg_b200._lib import EPI_BIAS_GELU, EPI_DGELU
    g = torch.Generator(device="cuda").manual_seed(M + N)
    a = (torch.randn(M, K, device=cuda, generator=g) * 0.5).bfloat16()
    b = (torch.randn(N, K, device=cuda, generator=g) * 0.1).bfloat16()
    bias = torch.randn(N, device=cuda, generator=g) * 0.1
    pre = a.float() @ b.float().T + bias
    out = ops.gemm(a, b, M=M, N=N, K=K, bias=bias, tile_n=tile_n, _dbg=mode)
    assert _rel(out, pre) < 6e-3
    legacy = ops.gemm(a, b, M=M, N=N, K=K, bias=bias, tile_n=tile_n, _dbg=mode | 8)          # direct-store epilogue
    assert torch.equal(out, legacy)
    aux = torch.full((M, N), float("nan"), device=cuda, dtype=torch.bfloat16)
    act = ops.gemm(a, b, M=M, N=N, K=K, bias=bias, epilogue=EPI_BIAS_GELU, aux=aux, tile_n=tile_n, _dbg=mode | 256)   # 256: 16-warp epilogue at any K
    assert torch.equal(aux, out)                                                  # same rounding of the pre-activation
    assert (act.float() - gelu(aux.float())).abs().max().item() < 2e-2
    act_noaux = ops.gemm(a, b, M=M, N=N, K=K, bias=bias, epilogue=EPI_BIAS_GELU, tile_n=tile_n, _dbg=mode)
    assert torch.equal(act_noaux, act)
    # backward: d = (a @ b^T) * gelu'(aux), aux2 = gelu(aux), col_sum += sum_m d
    cs = torch.zeros(N, device=cuda)
    a2 = torch.full((M, N), float("nan"), device=cuda, dtype=torch.bfloat16)
    dh = ops.gemm(a, b, M=M, N=N, K=K, epilogue=EPI_DGELU, aux=aux, aux2=a2, col_sum=cs, tile_n=tile_n, _dbg=mode)
    h = aux.float().requires_grad_(True)
    gelu(h).backward(pre - bias)
    assert _rel(dh, h.grad) < 6e-3
    assert _rel(cs, h.grad.sum(0)) < 2e-3
    assert (a2.float() - gelu(aux.float())).abs().max().item() < 2e-2 and torch.isfinite(a2.float()).all()
    assert (a2.float() - act.float()).abs().max().item() < 1e-6 + 8e-3 * act.float().abs().max().item()
    # 16-warp activation epilogues (default where the staging area fits) == the 8-warp version (reserved bit 64), bit for bit
    aux8 = torch.empty_like(aux)
    act8 = ops.gemm(a, b, M=M, N=N, K=K, bias=bias, epilogue=EPI_BIAS_GELU, aux=aux8, tile_n=tile_n, _dbg=mode | 64)
    assert torch.equal(act8, act) and torch.equal(aux8, aux)
    cs8 = torch.zeros(N, device=cuda)
    a28 = torch.empty_like(a2)
    dh8 = ops.gemm(a, b, M=M, N=N, K=K, epilogue=EPI_DGELU, aux=aux, aux2=a28, col_sum=cs8, tile_n=tile_n, _dbg=mode | 64)
    assert torch.equal(dh8, dh) and torch.equal(a28, a2)
    assert _rel(cs8, cs) < 1e-5
    # without the gelu(pre) re-emit (towers.keep_gelu_output: the forward's activation was kept): same gradient, same sums
    for extra in (0, 64):
        csn = torch.zeros(N, device=cuda)
        dhn = ops.gemm(a, b, M=M, N=N, K=K, epilogue=EPI_DGELU, aux=aux, col_sum=csn, tile_n=tile_n, _dbg=mode | extra)
        assert torch.equal(dhn, dh)
        assert _rel(csn, cs) < 1e-5


@pytest.mark.parametrize("D", [384, 768])
def test_add_layernorm(cuda, D):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(D)
    M = 777
    x = torch.randn(M, D, device=cuda, generator=g)
    add = torch.randn(M, D, device=cuda, generator=g).bfloat16()
    gamma = 1 + 0.1 * torch.randn(D, device=cuda, generator=g)
    beta = 0.1 * torch.randn(D, device=cuda, generator=g)
    s, yb, yf, mean, rstd = ops.add_layernorm_fwd(x, add, gamma, beta, 1e-6, want_f32=True)
    ref_s = x + add.float()
    ref = torch.nn.functional.layer_norm(ref_s, (D,), gamma, beta, 1e-6)
    assert torch.equal(s, ref_s)
    assert (yf - ref).abs().max().item() < 1e-4 and _rel(yb, ref) < 6e-3
    assert (mean - ref_s.mean(-1)).abs().max().item() < 1e-5
    assert _rel(rstd, (ref_s.var(-1, unbiased=False) + 1e-6).rsqrt()) < 1e-5


def test_gemm_tf32(cuda):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(5)
    M, N, K = 500, 700, 512
    a = torch.nn.functional.normalize(torch.randn(M, K, device=cuda, generator=g), dim=-1)
    b = torch.nn.functional.normalize(torch.randn(N, K, device=cuda, generator=g), dim=-1)
    out = ops.gemm(a, b, M=M, N=N, K=K, out_dtype=torch.float32)
    assert (out - a @ b.T).abs().max().item() < 1e-3
    # 32-bit MN-major operands are rejected loudly (they need the 128B_BASE32B smem layout)
    from simseg_b200._lib import SimsegError
    with pytest.raises(SimsegError):
        ops.gemm(a, b.T.contiguous(), M=M, N=N, K=K, b_major=1, out_dtype=torch.float32)


@pytest.mark.parametrize("D", [384, 768, 512])
@pytest.mark.parametrize("xdt", [torch.float32, torch.bfloat16])
def test_layernorm(cuda, D, xdt):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(D)
    M = 1003
    x = (torch.randn(M, D, device=cuda, generator=g) * 2 + 0.3).to(xdt)
    gamma = 1 + 0.1 * torch.randn(D, device=cuda, generator=g)
    beta = 0.1 * torch.randn(D, device=cuda, generator=g)
    eps = 1e-6
    yb, yf, mean, rstd = ops.layernorm_fwd(x, gamma, beta, eps, want_f32=True)
    xr = x.float().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(xr, (D,), gamma, beta, eps)
    assert (yf - ref).abs().max().item() < 2e-5
    assert _rel(yb, ref) < 6e-3
    dy = torch.randn(M, D, device=cuda, generator=g)
    dy2 = torch.randn(M, D, device=cuda, generator=g)
    gr = gamma.clone().requires_grad_(True); br = beta.clone().requires_grad_(True)
    torch.nn.functional.layer_norm(xr, (D,), gr, br, eps).backward(dy + dy2)
    prev = torch.randn(M, D, device=cuda, generator=g)
    dx = prev.clone()
    dxb = torch.empty(M, D, device=cuda, dtype=torch.bfloat16)
    dgam = torch.zeros(D, device=cuda); dbet = torch.zeros(D, device=cuda); dcs = torch.zeros(D, device=cuda)
    ops.layernorm_bwd(dy, x, gamma, mean, rstd, dy2=dy2, dx=dx, dx_accumulate=True, dx_bf16=dxb, dgamma=dgam, dbeta=dbet,
                      dx_colsum=dcs)
    assert (dx - (prev + xr.grad)).abs().max().item() < 1e-4
    assert _rel(dxb, prev + xr.grad) < 6e-3
    assert _rel(dgam, gr.grad) < 1e-4 and _rel(dbet, br.grad) < 1e-4
    assert _rel(dcs, (prev + xr.grad).sum(0)) < 1e-3


@pytest.mark.parametrize("B,H,S,masked", [(3, 6, 197, False), (2, 12, 325, False), (5, 12, 25, True), (4, 12, 77, True),
                                          (2, 2, 16, False), (2, 3, 64, True)])
def test_attention(cuda, B, H, S, masked):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(S)
    D = H * 64
    qkv = (torch.randn(B, S, 3, H, 64, device=cuda, generator=g)).bfloat16()
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    strides = (S * 3 * D, 3 * D, 64)
    klen = None
    if masked:
        klen = torch.randint(1, S + 1, (B,), device=cuda, generator=g, dtype=torch.int32)
        klen[0] = S
    scale = 0.125
    out, lse = ops.attention_fwd(q, k, v, B, H, S, strides, klen, scale)
    qf, kf, vf = [t.float().permute(0, 2, 1, 3).detach().requires_grad_(True) for t in (q, k, v)]
    s = (qf @ kf.transpose(-1, -2)) * scale
    if masked:
        km = torch.arange(S, device=cuda)[None] >= klen[:, None]
        s = s.masked_fill(km[:, None, None, :], float("-inf"))
    ref = (torch.softmax(s, -1) @ vf)
    ref_o = ref.permute(0, 2, 1, 3).reshape(B, S, D)
    assert (out.float() - ref_o).abs().max().item() < 3e-2
    assert (lse - torch.logsumexp(s, -1)).abs().max().item() < 2e-3
    dout = torch.randn(B, S, D, device=cuda, generator=g).bfloat16()
    ref_o.backward(dout.float())
    dqkv = torch.empty_like(qkv)
    ops.attention_bwd(q, k, v, out, dout, lse, B, H, S, strides, klen, scale, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2])
    for i, (name, t) in enumerate((("dq", qf), ("dk", kf), ("dv", vf))):
        r = t.grad.permute(0, 2, 1, 3)
        assert _rel(dqkv[:, :, i], r) < 2.5e-2, name


@pytest.mark.parametrize("impl", ["tc", "mma"])
@pytest.mark.parametrize("B,H,S,masked", [(3, 6, 197, False), (2, 12, 256, True), (7, 12, 25, True), (4, 12, 77, True),
                                          (2, 2, 129, True), (40, 6, 197, False), (3, 1, 128, False),
                                          (5, 8, 30, True), (3, 12, 50, False), (4, 6, 25, True), (200, 12, 25, True), (9, 16, 16, True)])
def test_attention_bwd_both_kernels(cuda, impl, B, H, S, masked):
    """Backward through the tcgen05 kernel (S <= 256) and through the mma.sync kernel, selected explicitly; more work
    items than SMs (40 x 6 heads) exercises the persistent loop, the buffer recycling and every barrier phase."""
    import os
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(S + B)
    D = H * 64
    qkv = (torch.randn(B, S, 3, H, 64, device=cuda, generator=g)).bfloat16()
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    strides = (S * 3 * D, 3 * D, 64)
    klen = None
    if masked:
        klen = torch.randint(1, S + 1, (B,), device=cuda, generator=g, dtype=torch.int32)
        klen[0] = S
    out, lse = ops.attention_fwd(q, k, v, B, H, S, strides, klen, 0.125)
    qf, kf, vf = [t.float().permute(0, 2, 1, 3).detach().requires_grad_(True) for t in (q, k, v)]
    s = (qf @ kf.transpose(-1, -2)) * 0.125
    if masked:
        km = torch.arange(S, device=cuda)[None] >= klen[:, None]
        s = s.masked_fill(km[:, None, None, :], float("-inf"))
    ref_o = (torch.softmax(s, -1) @ vf).permute(0, 2, 1, 3).reshape(B, S, D)
    dout = torch.randn(B, S, D, device=cuda, generator=g).bfloat16()
    ref_o.backward(dout.float())
    dqkv = torch.full_like(qkv, float("nan"))
    os.environ["SIMSEG_ATTN_BWD"] = impl
    try:
        ops.attention_bwd(q, k, v, out, dout, lse, B, H, S, strides, klen, 0.125, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2])
        torch.cuda.synchronize()
    finally:
        os.environ.pop("SIMSEG_ATTN_BWD", None)
    assert torch.isfinite(dqkv.float()).all()
    for i, (name, t) in enumerate((("dq", qf), ("dk", kf), ("dv", vf))):
        assert _rel(dqkv[:, :, i], t.grad.permute(0, 2, 1, 3)) < 2.5e-2, name


def test_embeddings(cuda):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(1)
    B, D = 3, 384
    img = torch.randn(B, 3, 224, 224, device=cuda, generator=g)
    w = torch.randn(D, 3, 16, 16, device=cuda, generator=g) * 0.02
    patches = ops.im2col16(img)
    ref = torch.nn.functional.conv2d(img.bfloat16().float(), w.bfloat16().float(), stride=16).flatten(2).transpose(1, 2)
    out = ops.gemm(patches, w.reshape(D, 768).bfloat16(), M=B * 196, N=D, K=768, out_dtype=torch.float32)
    assert _rel(out.reshape(B, 196, D), ref) < 1e-4
    cls = torch.randn(D, device=cuda, generator=g); pos = torch.randn(197, D, device=cuda, generator=g)
    x = ops.vit_tokens_fwd(out, cls, pos, B, 196, D)
    refx = torch.cat([cls.expand(B, 1, D), out.reshape(B, 196, D)], 1) + pos
    assert (x - refx).abs().max().item() < 1e-6
    dx = torch.randn(B, 197, D, device=cuda, generator=g)
    dpos = torch.zeros(197, D, device=cuda); dcls = torch.zeros(D, device=cuda)
    dpatch = ops.vit_tokens_bwd(dx, B, 196, D, dpos, dcls)
    assert _rel(dpatch.reshape(B, 196, D), dx[:, 1:]) < 5e-3
    assert (dpos - dx.sum(0)).abs().max().item() < 1e-5 and (dcls - dx[:, 0].sum(0)).abs().max().item() < 1e-5
    # BERT embeddings
    T, V, H = 25, 1000, 768
    ids = torch.randint(0, V, (B, T), device=cuda, generator=g); ids[:, 0] = 101
    word = torch.randn(V, H, device=cuda, generator=g); pe = torch.randn(512, H, device=cuda, generator=g)
    te = torch.randn(2, H, device=cuda, generator=g)
    e = ops.bert_embed_fwd(ids, word, pe, te)
    assert (e - (word[ids] + pe[:T] + te[0])).abs().max().item() < 1e-6
    de = torch.randn(B, T, H, device=cuda, generator=g)
    dword = torch.zeros_like(word); dpe = torch.zeros_like(pe); dte = torch.zeros_like(te)
    ops.bert_embed_bwd(ids, de, dword, dpe, dte[0])
    rw = torch.zeros_like(word).index_add_(0, ids.reshape(-1), de.reshape(-1, H))
    assert (dword - rw).abs().max().item() < 1e-5
    assert (dpe[:T] - de.sum(0)).abs().max().item() < 1e-5 and (dte[0] - de.sum((0, 1))).abs().max().item() < 1e-4


def test_misc_elementwise(cuda):
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(2)
    w = torch.randn(300, 77, device=cuda, generator=g)
    wb, wt = ops.cast_bf16(w, transpose_too=True)
    assert torch.equal(wb, w.bfloat16()) and torch.equal(wt, w.bfloat16().T.contiguous())
    x = torch.randn(5000, 1152, device=cuda, generator=g)
    assert _rel(ops.colsum(x), x.sum(0)) < 1e-5
    assert _rel(ops.colsum(x.bfloat16()), x.bfloat16().float().sum(0)) < 1e-5
    h = torch.randn(64, 1536, device=cuda, generator=g).bfloat16()
    assert (ops.gelu_fwd(h).float() - gelu(h.float())).abs().max().item() < 2e-2


@pytest.mark.parametrize("impl", ["tc", "mma", "ts"])
@pytest.mark.parametrize("B,H,S,masked", [(3, 6, 197, False), (2, 12, 224, True), (7, 12, 25, True), (4, 12, 77, True),
                                          (2, 2, 129, True), (60, 6, 197, False), (3, 1, 128, False), (300, 2, 64, True),
                                          (5, 8, 30, True), (3, 12, 50, False), (4, 6, 25, True), (200, 12, 25, True), (9, 16, 16, True)])
def test_attention_fwd_both_kernels(cuda, impl, B, H, S, masked, monkeypatch):
    """Forward through the tcgen05 kernel (S <= 224: exact two-pass softmax out of TMEM, two alternating softmax
    groups) and through the mma.sync kernel, selected explicitly; more work items than SMs exercises the persistent
    loop, the K/V / Q / S / P / O buffer recycling and every barrier phase."""
    ops = _ops()
    monkeypatch.setenv("SIMSEG_ATTN_FWD", impl)
    g = torch.Generator(device="cuda").manual_seed(S + B)
    D = H * 64
    qkv = (torch.randn(B, S, 3, H, 64, device=cuda, generator=g)).bfloat16()
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    strides = (S * 3 * D, 3 * D, 64)
    klen = None
    if masked:
        klen = torch.randint(1, S + 1, (B,), device=cuda, generator=g, dtype=torch.int32)
        klen[0] = S
    out = torch.full((B, S, D), float("nan"), device=cuda, dtype=torch.bfloat16)
    lse = torch.full((B, H, S), float("nan"), device=cuda)
    ops.attention_fwd(q, k, v, B, H, S, strides, klen, 0.125, out=out, lse=lse)
    qf, kf, vf = [t.float().permute(0, 2, 1, 3) for t in (q, k, v)]
    s = (qf @ kf.transpose(-1, -2)) * 0.125
    if masked:
        km = torch.arange(S, device=cuda)[None] >= klen[:, None]
        s = s.masked_fill(km[:, None, None, :], float("-inf"))
    ref_o = (torch.softmax(s, -1) @ vf).permute(0, 2, 1, 3).reshape(B, S, D)
    assert torch.isfinite(out.float()).all() and torch.isfinite(lse).all()
    assert (out.float() - ref_o).abs().max().item() < 3e-2
    assert (lse - torch.logsumexp(s, -1)).abs().max().item() < 2e-3


@pytest.mark.parametrize("T,Do,Di", [(30000, 384, 1536), (30000, 1536, 384), (20000, 1152, 384), (9000, 384, 384),
                                     (16000, 768, 3072), (16000, 2304, 768), (5000, 448, 320), (4100, 512, 171 * 0 + 192)])
def test_gemm_wgrad_wide_pair_tiles(cuda, T, Do, Di):
    """Split-K weight gradients on CTA-pair 256 x 384|512 tiles (one TMEM accumulator, two MMAs per k-step, dW or dW^T
    whichever pads less, reduction stores) against the single-CTA path (_dbg=128 disables the wide tiles) and fp32 torch."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(T + Do)
    dy = torch.randn(T, Do, device=cuda, generator=g).bfloat16()
    x = torch.randn(T, Di, device=cuda, generator=g).bfloat16()
    ref = dy.float().T @ x.float()
    out = torch.full((Do, Di), float("nan"), device=cuda)
    ops.linear_wgrad(dy, x, out)
    assert _rel(out, ref) < 1e-4
    ops.linear_wgrad(dy, x, out, accumulate=True)
    assert _rel(out, 2 * ref) < 1e-4
    legacy = ops.gemm(dy, x, M=Do, N=Di, K=T, a_major=1, b_major=1, out_dtype=torch.float32, _dbg=128)
    assert _rel(legacy, ref) < 1e-4


@pytest.mark.parametrize("B,H,S,masked", [(2, 12, 325, False), (64, 12, 325, False), (3, 6, 325, True), (2, 6, 384, True), (5, 6, 257, False),
                                          (4, 12, 256, True), (3, 6, 225, False), (200, 6, 325, False)])
def test_attention_fwd_long_sequences_on_tcgen05(cuda, B, H, S, masked, monkeypatch):
    """224 < S <= 384 — the geometry the reference's segmentation tool runs (288 x 288 -> 325 tokens,
    configs/clip/simseg.vit-s.yaml:70-77, tools/seg_evaluation.py:228-231) — through the tcgen05 forward (one S accumulator
    of up to 384 TMEM columns, two MMAs per S = Q K^T, three K / V tiles in smem) and bit-compared in value with the
    mma.sync kernel that used to take these shapes."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(S + B)
    D = H * 64
    qkv = (torch.randn(B, S, 3, H, 64, device=cuda, generator=g)).bfloat16()
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    strides = (S * 3 * D, 3 * D, 64)
    klen = None
    if masked:
        klen = torch.randint(1, S + 1, (B,), device=cuda, generator=g, dtype=torch.int32)
        klen[0] = S
    res = {}
    for impl in ("tc", "mma"):
        monkeypatch.setenv("SIMSEG_ATTN_FWD", impl)
        out = torch.full((B, S, D), float("nan"), device=cuda, dtype=torch.bfloat16)
        lse = torch.full((B, H, S), float("nan"), device=cuda)
        n0 = ops.launch_count()
        ops.attention_fwd(q, k, v, B, H, S, strides, klen, 0.125, out=out, lse=lse)
        assert ops.launch_count() == n0 + 1
        res[impl] = (out, lse)
    qf, kf, vf = [t.float().permute(0, 2, 1, 3) for t in (q, k, v)]
    s = (qf @ kf.transpose(-1, -2)) * 0.125
    if masked:
        km = torch.arange(S, device=cuda)[None] >= klen[:, None]
        s = s.masked_fill(km[:, None, None, :], float("-inf"))
    ref_o = (torch.softmax(s, -1) @ vf).permute(0, 2, 1, 3).reshape(B, S, D)
    for impl, (out, lse) in res.items():
        assert torch.isfinite(out.float()).all() and torch.isfinite(lse).all(), impl
        assert (out.float() - ref_o).abs().max().item() < 3e-2, impl
        assert (lse - torch.logsumexp(s, -1)).abs().max().item() < 2e-3, impl
    assert (res["tc"][0].float() - res["mma"][0].float()).abs().max().item() < 2e-2
