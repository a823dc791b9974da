"""Train-mode dropout of the BERT tower (HF BertModel under model.train(): hidden_dropout_prob = attention_probs_dropout_prob
= 0.1; the reference trains it so through ``huggingface_builder.py:16-17``).

The product's masks are a counter-based Philox stream (csrc/philox.cuh); the oracle restates the stream in numpy
(known-answer tested on the CPU) and its dropout PLACEMENT is pinned against transformers' BertModel.train()
(tests/golden/bert_dropout.npz).  Here: the kernels' keep bits equal the oracle's bit for bit, every fused dropout
(LayerNorm epilogues, attention probabilities) matches plain torch math given those bits in forward and backward, and the
whole tower in train mode matches the oracle with the same {seed, step}.
"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

SEED = 0x1234_5678_9ABC


def _ops():
    from simseg_b200 import ops
    return ops


def _rel(a, b):
    return ((a.float() - b.float()).abs().max() / (b.float().abs().max() + 1e-12)).item()


def _rng(cuda, seed=SEED, step=5):
    return torch.tensor([seed, step], dtype=torch.int64, device=cuda)


def _pack_factor(H, S):
    G = 1
    while G * 2 * S <= 128 and H % (G * 2) == 0 and G < 8:
        G *= 2
    return G


def _decode_attn_mask(words, B, H, S):
    """int32 words in the kernels' tile coordinates (include/simseg_b200.h:simseg_attn_dropout_mask) -> bool [B,H,S,S]."""
    G = _pack_factor(H, S)
    rows, nw = S * G, (S * G + 31) // 32
    w = words.cpu().numpy().view(np.uint32).reshape(B, H // G, rows, nw)
    bits = ((w[..., None] >> np.arange(32, dtype=np.uint32)) & 1).reshape(B, H // G, rows, nw * 32)[..., :rows].astype(bool)
    keep = np.zeros((B, H, S, S), dtype=bool)
    for hg in range(H // G):
        for g in range(G):
            # tile row r = q * G + g, tile column c = k * G + g
            keep[:, hg * G + g] = bits[:, hg, g::G, g::G]
    return torch.from_numpy(keep)


@pytest.mark.parametrize("B,H,S", [(5, 12, 25), (3, 12, 40), (2, 12, 77), (2, 6, 197), (3, 8, 16), (2, 12, 7)])
def test_attention_keep_bits_equal_the_oracle_stream(cuda, B, H, S):
    """Packed (G = 4, 2, 8) and unpacked tiles: every live bit equals keep(b,h,q,k) of the numpy Philox restatement."""
    from oracle import simseg_oracle as O
    ops = _ops()
    words = ops.attn_dropout_mask(B, H, S, ops.Drop(0.1, _rng(cuda), 7))
    torch.cuda.synchronize()
    got = _decode_attn_mask(words, B, H, S)
    want = O.attn_keep_mask(SEED, 5, 7, B, H, S, 0.1)
    assert torch.equal(got, want)
    assert abs(got.float().mean().item() - 0.9) < 0.02


@pytest.mark.parametrize("D", [768, 384])
def test_layernorm_dropout_forward_backward_vs_oracle_mask(cuda, D):
    """Both fused forms: LayerNorm(x + dropout(add)) (BertSelfOutput / BertOutput) and dropout(LayerNorm(x)) (BertEmbeddings),
    forward and backward, against fp32 torch math with the oracle's keep bits."""
    from oracle import simseg_oracle as O
    ops = _ops()
    M, p = 333, 0.1
    g = torch.Generator(device="cuda").manual_seed(D)
    x = torch.randn(M, D, device=cuda, generator=g)
    add = torch.randn(M, D, device=cuda, generator=g).bfloat16()
    gam = torch.randn(D, device=cuda, generator=g) * 0.2 + 1
    bet = torch.randn(D, device=cuda, generator=g) * 0.1
    rng = _rng(cuda, step=9)
    keep = O.hidden_keep_mask(SEED, 9, 3, M, D, p).to(cuda).float() / (1 - p)
    # ---- form 1
    s, yb, yf, mean, rstd = ops.add_layernorm_fwd(x, add, gam, bet, 1e-12, want_f32=True, drop=ops.Drop(p, rng, 3))
    xr, ar = x.clone().requires_grad_(True), add.float().requires_grad_(True)
    gr, br = gam.clone().requires_grad_(True), bet.clone().requires_grad_(True)
    s_ref = xr + ar * keep
    y_ref = torch.nn.functional.layer_norm(s_ref, (D,), gr, br, 1e-12)
    assert (s - s_ref).abs().max().item() < 1e-5
    assert (s == x).eq(keep == 0).float().mean().item() > 0.999          # dropped elements contribute exactly nothing
    assert (yf - y_ref).abs().max().item() < 1e-4 and _rel(yb, y_ref) < 6e-3
    dy = torch.randn(M, D, device=cuda, generator=g)
    dy2 = torch.randn(M, D, device=cuda, generator=g)
    y_ref.backward(dy + dy2)
    dx = torch.empty(M, D, device=cuda)
    dxb = torch.empty(M, D, device=cuda, dtype=torch.bfloat16)
    dgam, dbet, dcs = (torch.zeros(D, device=cuda) for _ in range(3))
    ops.layernorm_bwd(dy, s, gam, mean, rstd, dy2=dy2, dx=dx, dx_bf16=dxb, dgamma=dgam, dbeta=dbet, dx_colsum=dcs,
                      drop=ops.Drop(p, rng, 3), drop_mode=1)
    torch.cuda.synchronize()
    assert _rel(dx, xr.grad) < 1e-4                                      # residual path: unmasked
    assert _rel(dxb, ar.grad) < 6e-3                                     # dense path: masked and scaled
    assert ((dxb.float() == 0) | (keep > 0)).all()
    assert _rel(dcs, ar.grad.sum(0)) < 1e-3
    assert _rel(dgam, gr.grad) < 1e-4 and _rel(dbet, br.grad) < 1e-4
    # ---- form 2
    yb2, yf2, mean2, rstd2 = ops.layernorm_fwd(x, gam, bet, 1e-12, want_f32=True, drop=ops.Drop(p, rng, 3))
    xr = x.clone().requires_grad_(True)
    gr, br = gam.clone().requires_grad_(True), bet.clone().requires_grad_(True)
    y2_ref = torch.nn.functional.layer_norm(xr, (D,), gr, br, 1e-12) * keep
    assert (yf2 - y2_ref).abs().max().item() < 1e-4 and _rel(yb2, y2_ref) < 6e-3
    y2_ref.backward(dy + dy2)
    dx2 = torch.empty(M, D, device=cuda)
    dgam.zero_(); dbet.zero_()
    ops.layernorm_bwd(dy, x, gam, mean2, rstd2, dy2=dy2, dx=dx2, dgamma=dgam, dbeta=dbet, drop=ops.Drop(p, rng, 3), drop_mode=2)
    torch.cuda.synchronize()
    assert _rel(dx2, xr.grad) < 1e-4
    assert _rel(dgam, gr.grad) < 1e-4 and _rel(dbet, br.grad) < 1e-4
    # another site / step -> another mask; p = 0 entry == the plain kernel
    s_b = ops.add_layernorm_fwd(x, add, gam, bet, 1e-12, drop=ops.Drop(p, rng, 4))[0]
    assert not torch.equal(s_b, s)
    rng[1] += 1
    assert not torch.equal(ops.add_layernorm_fwd(x, add, gam, bet, 1e-12, drop=ops.Drop(p, rng, 3))[0], s)


@pytest.mark.parametrize("B,H,S,masked", [(7, 12, 25, True), (4, 12, 77, True), (3, 12, 40, True), (3, 6, 197, False),
                                          (200, 12, 25, True), (2, 12, 128, True)])
def test_attention_dropout_forward_backward_vs_torch(cuda, B, H, S, masked):
    """out = dropout(softmax(q k^T / 8 + mask)) v with the kernel's own keep bits (checked against the oracle stream above):
    output, log-sum-exp (of the UNdropped softmax) and dq / dk / dv against fp32 torch autograd."""
    ops = _ops()
    p = 0.1
    g = torch.Generator(device="cuda").manual_seed(S + B)
    D = H * 64
    qkv = torch.randn(B, S, 3, H, 64, device=cuda, generator=g).bfloat16()
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    strides = (S * 3 * D, 3 * D, 64)
    klen = None
    if masked:
        klen = torch.randint(1, S + 1, (B,), device=cuda, generator=g, dtype=torch.int32)
        klen[0] = S
    words = ops.attn_dropout_mask(B, H, S, ops.Drop(p, _rng(cuda), 1))
    keep = _decode_attn_mask(words, B, H, S).to(cuda).float() / (1 - p)
    out, lse = ops.attention_fwd(q, k, v, B, H, S, strides, klen, 0.125, drop_mask=words, drop_p=p)
    qf, kf, vf = [t.float().permute(0, 2, 1, 3).detach().requires_grad_(True) for t in (q, k, v)]
    s = (qf @ kf.transpose(-1, -2)) * 0.125
    if masked:
        km = torch.arange(S, device=cuda)[None] >= klen[:, None]
        s = s.masked_fill(km[:, None, None, :], float("-inf"))
    ref_o = ((torch.softmax(s, -1) * keep) @ vf).permute(0, 2, 1, 3).reshape(B, S, D)
    assert (out.float() - ref_o).abs().max().item() < 3e-2
    assert (lse - torch.logsumexp(s, -1)).abs().max().item() < 2e-3
    plain, _ = ops.attention_fwd(q, k, v, B, H, S, strides, klen, 0.125)
    assert (plain.float() - out.float()).abs().max().item() > 0.05           # the mask acts
    dout = torch.randn(B, S, D, device=cuda, generator=g).bfloat16()
    ref_o.backward(dout.float())
    dqkv = torch.full_like(qkv, float("nan"))
    ops.attention_bwd(q, k, v, out, dout, lse, B, H, S, strides, klen, 0.125, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2],
                      drop_mask=words, drop_p=p)
    torch.cuda.synchronize()
    assert torch.isfinite(dqkv.float()).all()
    for i, (name, t) in enumerate((("dq", qf), ("dk", kf), ("dv", vf))):
        assert _rel(dqkv[:, :, i], t.grad.permute(0, 2, 1, 3)) < 2.5e-2, name


def test_attention_dropout_unsupported_shapes_fail_loudly(cuda):
    from simseg_b200._lib import SimsegError
    ops = _ops()
    with pytest.raises(SimsegError):
        ops.attn_dropout_mask(2, 12, 300, ops.Drop(0.1, _rng(cuda), 1))


def _build(cuda):
    from simseg_b200.config import load_cfg
    from simseg_b200.pipeline import PIPELINE
    cfg = load_cfg("simseg.vit-s.yaml", ["model.image_encoder.pretrained=False", "model.text_encoder.pretrained=False",
                                          "transforms.input_size=224"])
    return PIPELINE["clip"](cfg).to(cuda), cfg


def _cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return (a @ b / (a.norm() * b.norm() + 1e-30)).item()


def test_constructor_defaults_are_the_reference_probabilities(cuda, monkeypatch):
    monkeypatch.delenv("SIMSEG_BERT_DROPOUT", raising=False)
    model, _ = _build(cuda)
    hf = model.text_encoder.model
    assert hf.hidden_dropout_prob == 0.1 and hf.attention_probs_dropout_prob == 0.1      # bert-base-uncased config.json
    assert "drop_rng" not in "".join(model.state_dict().keys())                          # state-dict surface unchanged


@pytest.mark.parametrize("T", [25, 77])
def test_bert_tower_train_mode_vs_oracle_same_stream(cuda, T):
    """Text tower under model.train() with p = 0.1 vs the fp32 oracle fed the SAME {seed, step}: tokens, and every BERT
    parameter gradient for a fixed cotangent.  eval() on the same model reproduces the p = 0 oracle."""
    from oracle import simseg_oracle as O
    model, _ = _build(cuda)
    hf = model.text_encoder.model
    hf.hidden_dropout_prob = hf.attention_probs_dropout_prob = 0.1
    sd = O.make_state_dict(384, 6, seed=0)
    model.load_state_dict(sd, strict=True)
    batch = O.make_batch(6, T, seed=77)
    ids, am = batch["input_ids"].to(cuda), batch["attention_mask"].to(cuda)
    model.train()
    hf.seed_dropout(SEED, 41)                       # the forward below runs at step 42
    model.zero_grad(set_to_none=True)
    tok = model.forward_text_feature(ids, am)
    cot = torch.randn(tok.shape, generator=torch.Generator().manual_seed(3)).to(cuda) * am[..., None]
    tok.backward(cot)
    torch.cuda.synchronize()
    assert int(hf.drop_rng[1].item()) == 42
    sdg = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items() if k.startswith(O.TXT_PREFIX)}
    ref = O.bert_forward(sdg, batch["input_ids"], batch["attention_mask"], 12, O.TXT_PREFIX,
                         dropout=O.PhiloxDropout(SEED, 42, 0.1, 0.1))
    (ref * cot.cpu()).sum().backward()
    live = batch["attention_mask"].bool()
    assert (tok.cpu()[live] - ref.detach()[live]).abs().max().item() < 0.12            # bf16 GEMM operands, 12 layers
    assert _cos(tok.cpu()[live], ref.detach()[live]) > 0.999
    with torch.no_grad():
        ref0 = O.bert_forward(sd, batch["input_ids"], batch["attention_mask"], 12, O.TXT_PREFIX)
    assert _cos(ref0[live], ref.detach()[live]) < 0.99                                  # a different function of the input
    rows = []
    for k, p in model.named_parameters():
        if not k.startswith(O.TXT_PREFIX) or sdg[k].grad is None or sdg[k].grad.norm().item() <= 1e-7:
            continue
        if k.endswith("attention.self.key.bias"):       # exactly zero in theory (rows of dS sum to 0, dropout or not): noise / noise
            continue
        gq, gr = p.grad.cpu(), sdg[k].grad
        rows.append((k, _cos(gq, gr), ((gq - gr).norm() / gr.norm()).item()))
    bad = sorted(rows, key=lambda r: r[1])[:6]
    assert len(rows) > 180
    assert all(r[1] > 0.98 and r[2] < 0.2 for r in rows), bad
    # same seed / step again -> bit-identical tokens; next step -> different masks
    hf.seed_dropout(SEED, 41)
    with torch.no_grad():
        again = model.forward_text_feature(ids, am)
        nxt = model.forward_text_feature(ids, am)
    assert torch.equal(again, tok.detach()) and not torch.equal(nxt, again)
    model.eval()
    with torch.no_grad():
        ev = model.forward_text_feature(ids, am)
    assert _cos(ev.cpu()[live], ref0[live]) > 0.999
    assert int(hf.drop_rng[1].item()) == 43                                             # eval forwards do not advance the stream


def test_graph_replay_draws_fresh_masks_and_micro_batches_reuse_them(cuda):
    """{seed, step} lives on the device: a captured train step advances it on every replay (two replays of the same batch give
    different losses under dropout, equal ones at p = 0), and the two passes of the micro-batched step see the same steps."""
    from oracle import simseg_oracle as O
    from simseg_b200.train import Trainer
    model, cfg = _build(cuda)
    hf = model.text_encoder.model
    hf.hidden_dropout_prob = hf.attention_probs_dropout_prob = 0.1
    model.load_state_dict(O.make_state_dict(384, 6, seed=0), strict=True)
    model.train()
    hf.seed_dropout(SEED, 0)
    batch = {k: v.to(cuda) for k, v in O.make_batch(8, 25, seed=4).items()}
    tr = Trainer(model, cfg, micro_batch=4)
    step0 = int(hf.drop_rng[1].item())
    tr.backward_only(batch)
    torch.cuda.synchronize()
    assert int(hf.drop_rng[1].item()) == step0 + 2          # two micro-batches; pass 2 re-used the steps of pass 1
    tr2 = Trainer(model, cfg, capturable=True)
    for g in tr2.opt.param_groups:
        g["lr"].zero_()                                     # weights stay put: only the masks can change the loss
    gs = tr2.capture(batch, warmup=2)
    s0 = int(hf.drop_rng[1].item())
    l1 = gs(batch)[0].item()
    l2 = gs(batch)[0].item()
    assert int(hf.drop_rng[1].item()) == s0 + 2 and l1 != l2
