"""Parity of the head / loss / similarity kernels against the CPU oracle and the committed golden
fixtures (tests/golden/*.npz, generated from the reference's own modules by oracle/make_golden.py).

Tolerances (north_star): logits and similarity maps within 1e-3 for fp32 inputs, 1e-2 for bf16 inputs;
argmax masks bit-exact (checked where the oracle's top-1/top-2 margin exceeds 2x the tolerance).
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _ops():
    from simseg_b200 import ops
    return ops


def _O():
    from oracle import simseg_oracle as O
    return O


def test_topk_pool_l2norm_vs_golden(cuda):
    ops, O = _ops(), _O()
    gold = np.load(os.path.join(GOLD, "heads_loss.npz"))
    g = torch.Generator().manual_seed(int(gold["heads_seed"]))
    x_img = torch.randn(6, 196, 384, generator=g)
    x_txt = torch.randn(6, 25, 768, generator=g)
    assert np.array_equal(x_img[:2, :, :64].numpy(), gold["heads_x_img"])          # same synthetic inputs
    wi, wt = torch.tensor(gold["heads_wi"]), torch.tensor(gold["heads_wt"])
    mask = torch.tensor(gold["heads_mask"])
    # projection (exact fp32 product here: isolates the pooling kernel), then fused pool + L2norm
    pi = (x_img @ wi.T).to(cuda).contiguous()
    pt = (x_txt @ wt.T).to(cuda).contiguous()
    _, emb_i, idx_i = ops.topk_pool_l2norm_fwd(pi, 5, 0, 196)
    _, emb_t, idx_t = ops.topk_pool_l2norm_fwd(pt, 1, 0, 25, attention_mask=mask.to(cuda))
    assert np.abs(emb_i.cpu().numpy() - gold["heads_img_emb"]).max() < 1e-5
    assert np.abs(emb_t.cpu().numpy() - gold["heads_txt_emb"]).max() < 1e-5
    # backward vs autograd through the oracle
    pr = pi.cpu().clone().requires_grad_(True)
    d = torch.randn(6, 512, generator=g)
    O.l2norm(O.topk_pooling(pr, 5)).backward(d)
    pooled, emb, idx = ops.topk_pool_l2norm_fwd(pi, 5, 0, 196)
    dx = ops.topk_pool_l2norm_bwd(d.to(cuda), pooled, idx, 196, 5)
    ref = pr.grad
    assert ((dx.float().cpu() - ref).abs().max() / ref.abs().max()).item() < 6e-3
    assert torch.equal(dx.float().cpu() != 0, ref != 0)                              # same selected tokens


def test_topk_pool_with_cls_offset_bf16(cuda):
    ops, O = _ops(), _O()
    g = torch.Generator().manual_seed(3)
    x = torch.randn(4, 197, 512, generator=g).bfloat16()
    pooled, emb, _ = ops.topk_pool_l2norm_fwd(x.to(cuda), 5, 1, 196)
    ref = O.l2norm(O.topk_pooling(x.float()[:, 1:], 5))
    assert (emb.cpu() - ref).abs().max().item() < 1e-5


def _head_case(cuda, B, S, D, E, k, t0, masked, seed):
    g = torch.Generator().manual_seed(seed)
    x = (torch.randn(B, S, D, generator=g) * 0.7).bfloat16().to(cuda).contiguous()
    w = (torch.randn(E, D, generator=g) * 0.05).bfloat16().to(cuda).contiguous()
    mask = None
    if masked:
        lens = torch.randint(max(k, 1) + t0, S + 1, (B,), generator=g)
        lens[0] = S
        mask = (torch.arange(S)[None, :] < lens[:, None]).long().to(cuda).contiguous()
    demb = torch.randn(B, E, generator=g).to(cuda)
    return x, w, mask, demb


@pytest.mark.parametrize("B,S,D,E,k,t0,masked", [(5, 197, 384, 512, 5, 1, False),       # ViT-S image head (clip.py:87-93)
                                                  (3, 197, 768, 512, 5, 1, False),       # ViT-B
                                                  (23, 25, 768, 512, 1, 1, True),        # text head, T = 25, k = 1, masks
                                                  (7, 77, 768, 512, 3, 1, True),         # T = 77, clamped k
                                                  (9, 25, 768, 256, 8, 0, False),        # k = 8 (two samples per tile)
                                                  (301, 197, 384, 512, 5, 1, False)])    # more work items than SMs
def test_fused_projection_topk_head_equals_two_kernel_path(cuda, B, S, D, E, k, t0, masked):
    """f1: projection GEMM with the top-k pooling as its epilogue, and the backward with the dY operand generated in shared
    memory, against the two-kernel path (tcgen05 GEMM -> bf16 [B,S,E] -> ``topk_pool_l2norm`` / dense dgrad + wgrad) and the
    oracle (``pooling.py:57-65``, ``normalization.py:6-11``) on the same bf16 inputs."""
    ops, O = _ops(), _O()
    x, w, mask, demb = _head_case(cuda, B, S, D, E, k, t0, masked, 100 + B + S)
    ntok = S - t0
    # two-kernel path
    proj = ops.linear_fwd(x.view(B * S, D), w).view(B, S, E)
    pooled0, emb0, idx0 = ops.topk_pool_l2norm_fwd(proj, k, t0, ntok, attention_mask=mask)
    # fused
    pooled1, emb1, idx1 = ops.proj_topk_fwd(x, w, k, t0, ntok, attention_mask=mask)
    assert (pooled1 - pooled0).abs().max().item() <= 1e-6 + 4e-3 * pooled0.abs().max().item()
    assert (emb1 - emb0).abs().max().item() < 2e-3
    # selected tokens: the k largest bf16 projections, earliest token first among equal values (a stable descending sort) —
    # exactly, in both paths (the two GEMM tile shapes give bit-identical fp32-accumulated products)
    pm = proj.float().clone()
    if mask is not None:
        pm[mask == 0] = -10000.0
    stable = pm[:, t0:].sort(dim=1, descending=True, stable=True).indices[:, :k] + t0
    assert torch.equal(idx1.long(), stable)
    assert torch.equal(idx0.long(), stable)
    # oracle on the fp32 product of the same bf16 operands
    pf = (x.float().cpu() @ w.float().cpu().T)
    if mask is not None:
        pf = pf.clone()
        pf[mask.cpu() == 0] = -10000.0
    ref = O.l2norm(O.topk_pooling(pf[:, t0:], k))
    assert (emb1.cpu() - ref).abs().max().item() < 1e-2                       # bf16 rounding of the projection (north_star 1e-2)
    assert idx1.min().item() >= t0 and idx1.max().item() < S

    # ---- backward from the FUSED forward's own selection (so both paths see the same sparse dY)
    dy = ops.topk_pool_l2norm_bwd(demb, pooled1, idx1, S, k)                   # dense bf16 [B,S,E]
    wt = w.t().contiguous()
    dx0 = ops.linear_dgrad(dy.view(B * S, E), wt, out_dtype=torch.float32).view(B, S, D)
    dw0 = torch.zeros(E, D, device=cuda)
    ops.linear_wgrad(dy.view(B * S, E), x.view(B * S, D), dw0, accumulate=False)
    dw1 = torch.full((E, D), 0.25, device=cuda)                                # accumulates
    dx1 = ops.proj_topk_bwd(demb, pooled1, idx1, x, wt, k, dw=dw1)
    assert (dx1 - dx0).abs().max().item() <= 1e-6 + 2e-3 * dx0.abs().max().item()
    assert (dx1[:, :t0] == 0).all()
    assert ((dw1 - 0.25) - dw0).abs().max().item() <= 1e-6 + 2e-3 * dw0.abs().max().item()
    # and against fp32 autograd through the oracle: which of several bf16-EQUAL projections carries the gradient is a
    # tie-break (torch.topk's differs from "earliest token"), but the gradient summed over the tokens of a sample is not:
    # sum_s dx[b, s, :] = dL/dpooled[b, :] @ W
    pl = pooled1.cpu().clone().requires_grad_(True)
    O.l2norm(pl).backward(demb.cpu())
    dx_sum_ref = pl.grad @ w.float().cpu()
    assert (dx1.sum(1).cpu() - dx_sum_ref).abs().max().item() < 1e-2 * dx_sum_ref.abs().max().item() + 1e-6


def test_fused_head_vs_reference_fixture(cuda):
    """The fused head on the reference's own fixture (``tests/golden/heads_loss.npz``: SimpleProjection + TopKPooling + L2norm
    of ``clip.py:87-93,111-120`` run by the reference's modules), bf16 operands -> 1e-2."""
    ops = _ops()
    gold = np.load(os.path.join(GOLD, "heads_loss.npz"))
    g = torch.Generator().manual_seed(int(gold["heads_seed"]))
    x_img = torch.randn(6, 196, 384, generator=g)
    x_txt = torch.randn(6, 25, 768, generator=g)
    wi, wt = torch.tensor(gold["heads_wi"]), torch.tensor(gold["heads_wt"])
    mask = torch.tensor(gold["heads_mask"])
    _, emb_i, _ = ops.proj_topk_fwd(x_img.bfloat16().to(cuda).contiguous(), wi.bfloat16().to(cuda).contiguous(), 5, 0, 196)
    _, emb_t, _ = ops.proj_topk_fwd(x_txt.bfloat16().to(cuda).contiguous(), wt.bfloat16().to(cuda).contiguous(), 1, 0, 25,
                                    attention_mask=mask.to(cuda))
    assert np.abs(emb_i.cpu().numpy() - gold["heads_img_emb"]).max() < 1e-2
    assert np.abs(emb_t.cpu().numpy() - gold["heads_txt_emb"]).max() < 1e-2


@pytest.mark.parametrize("prec", [0, 1])
def test_infonce_vs_golden_and_oracle(cuda, prec):
    ops, O = _ops(), _O()
    gold = np.load(os.path.join(GOLD, "heads_loss.npz"))
    img = torch.tensor(gold["heads_img_emb"]); txt = torch.tensor(gold["heads_txt_emb"])
    temp = torch.tensor(0.02)
    b = img.shape[0]
    # tcgen05 kind::tf32 TRUNCATES fp32 operands to 10 mantissa bits: on these highly correlated embeddings
    # (cos ~ 0.96) the bias does not cancel, giving ~7e-4 relative on the cosine -> 4e-2 on logits of ~50
    tol = 1e-3 if prec == 0 else 5e-2
    gi, gt, gtemp = img.to(cuda), txt.to(cuda), temp.to(cuda)
    l1, lse1, am1, cos1, logits1 = ops.infonce_fwd(gi, gt, gtemp, 0, prec, want_logits=True)
    l2, lse2, am2, cos2, _ = ops.infonce_fwd(gt, gi, gtemp, 0, prec)
    loss = 0.5 * (l1.mean() + l2.mean())
    _, _, ref_logits, _ = O.nce_direction(img, txt, temp, 0)
    assert (logits1.cpu() - ref_logits).abs().max().item() < tol
    assert abs(loss.item() - float(gold["nce_loss"])) < tol
    assert abs((am1.cpu() == torch.arange(b)).float().mean().item() - float(gold["nce_i2t"])) < 1e-6
    assert abs((am2.cpu() == torch.arange(b)).float().mean().item() - float(gold["nce_t2i"])) < 1e-6
    # backward: d loss / d img, d txt, d temperature (reference autograd values in the fixture)
    dimg_g = torch.zeros_like(gi); dtxt_g = torch.zeros_like(gt); dtemp = torch.zeros((), device=cuda)
    d_img = ops.infonce_bwd(gi, gt, gtemp, 0, lse1, 0.5 / b, cos1, dtxt_g, dtemp, prec)
    d_txt = ops.infonce_bwd(gt, gi, gtemp, 0, lse2, 0.5 / b, cos2, dimg_g, dtemp, prec)
    rtol = 1e-4 if prec == 0 else 5e-3
    for got, ref in ((d_img + dimg_g, gold["nce_dimg"]), (d_txt + dtxt_g, gold["nce_dtxt"])):
        ref = torch.tensor(ref)
        assert ((got.cpu() - ref).abs().max() / ref.abs().max()).item() < rtol
    assert abs(dtemp.item() - float(gold["nce_dtemp"])) / abs(float(gold["nce_dtemp"])) < max(rtol, 1e-3)


def test_infonce_global_reduce_two_rank_fixture(cuda):
    """Per-rank losses / gathered-gradient sums of the reference's GatherLayer path (2 gloo ranks, fixture)."""
    ops = _ops()
    gold = np.load(os.path.join(GOLD, "global_reduce.npz"))
    img = torch.tensor(gold["gr_img"]).to(cuda); txt = torch.tensor(gold["gr_txt"]).to(cuda)
    temp = torch.tensor(0.02, device=cuda)
    W, b = 2, 6
    dimg_g = torch.zeros_like(img); dtxt_g = torch.zeros_like(txt)
    dloc_i, dloc_t = [], []
    for r in range(W):
        li, lt = img[r * b:(r + 1) * b].contiguous(), txt[r * b:(r + 1) * b].contiguous()
        l1, lse1, _, cos1, _ = ops.infonce_fwd(li, txt, temp, r * b)
        l2, lse2, _, cos2, _ = ops.infonce_fwd(lt, img, temp, r * b)
        loss = 0.5 * (l1.mean() + l2.mean())
        assert abs(loss.item() - float(gold[f"gr_loss_{r}"])) < 1e-4
        dloc_i.append(ops.infonce_bwd(li, txt, temp, r * b, lse1, 0.5 / b, cos1, dtxt_g, None))
        dloc_t.append(ops.infonce_bwd(lt, img, temp, r * b, lse2, 0.5 / b, cos2, dimg_g, None))
    for r in range(W):      # local-row gradient + this rank's slice of the all-reduced gathered gradient
        gi = dloc_i[r] + dimg_g[r * b:(r + 1) * b]
        gt = dloc_t[r] + dtxt_g[r * b:(r + 1) * b]
        assert (gi.cpu() - torch.tensor(gold[f"gr_dimg_{r}"])).abs().max().item() < 1e-5
        assert (gt.cpu() - torch.tensor(gold[f"gr_dtxt_{r}"])).abs().max().item() < 1e-5


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 1e-3), (torch.bfloat16, 1e-2)])
def test_patch_text_sim_golden(cuda, dtype, tol):
    ops = _ops()
    gold = np.load(os.path.join(GOLD, "heads_loss.npz"))
    p = torch.tensor(gold["seg_patch"]).to(cuda).to(dtype).contiguous()
    t = torch.tensor(gold["seg_text"]).to(cuda).to(dtype).contiguous()
    sim, am = ops.patch_text_sim(p, t)
    ref = torch.tensor(gold["seg_sim"])
    assert (sim.cpu() - ref).abs().max().item() < tol
    top2 = ref.topk(2, -1)[0]
    safe = (top2[..., 0] - top2[..., 1]) > 2 * tol
    assert safe.float().mean().item() > 0.2
    assert torch.equal(am.cpu().long()[safe], torch.tensor(gold["seg_argmax"])[safe])
    # argmax is bit-exact w.r.t. the map the kernel itself wrote
    assert torch.equal(am.long(), sim.argmax(-1))


@pytest.mark.parametrize("B,N,C,dtype", [(32, 196, 20, torch.bfloat16), (64, 196, 171, torch.bfloat16),
                                         (3, 324, 81, torch.float32), (1, 7, 1, torch.float32)])
def test_patch_text_sim_shapes_vs_oracle(cuda, B, N, C, dtype):
    ops, O = _ops(), _O()
    g = torch.Generator().manual_seed(B * N + C)
    p = torch.randn(B, N, 512, generator=g).to(dtype)
    t = torch.nn.functional.normalize(torch.randn(C, 512, generator=g), dim=-1).to(dtype)
    sim, am = ops.patch_text_sim(p.to(cuda), t.to(cuda))
    ref, _ = O.patch_text_sim(p, t)
    tol = 1e-3 if dtype == torch.float32 else 1e-2
    assert (sim.cpu() - ref).abs().max().item() < tol
    assert torch.equal(am.long(), sim.argmax(-1))


@pytest.mark.parametrize("rows,C,E", [(588, 171, 512), (1, 1, 512), (129, 255, 512), (4097, 256, 512), (300, 16, 256),
                                      (77, 300, 512), (640, 33, 1024)])
def test_patch_text_sim_fused_ragged_bf16(cuda, rows, C, E):
    """Edge shapes of the single-pass fused kernel (ragged last tile, C = 1 / 255 / 256, other E) and of the generic
    sequence it defers to (C > 256); an all-zero patch row must give similarity 0 and argmax 0 (F.normalize eps rule)."""
    ops, O = _ops(), _O()
    g = torch.Generator().manual_seed(rows * 7 + C)
    p = torch.randn(rows, E, generator=g).bfloat16()
    p[rows // 2] = 0
    t = torch.nn.functional.normalize(torch.randn(C, E, generator=g), dim=-1).bfloat16()
    sim, am = ops.patch_text_sim(p.to(cuda), t.to(cuda))
    ref, _ = O.patch_text_sim(p[None], t)
    assert sim.shape == (rows, C) and am.shape == (rows,)
    assert (sim.cpu() - ref[0]).abs().max().item() < 1e-2
    assert torch.equal(am.long(), sim.argmax(-1))
    assert float(sim[rows // 2].abs().max()) == 0.0 and int(am[rows // 2]) == 0
    # normalize=False is the plain projection product
    raw, _ = ops.patch_text_sim(p.to(cuda), t.to(cuda), normalize=False, want_argmax=False)
    assert (raw.cpu() - p.float() @ t.float().T).abs().max().item() < 0.15


def test_patch_text_sim_large_batch_properties(cuda):
    """Full-size property check (4096 maps x 196 patches x 171 classes — far beyond what the CPU oracle is run on):
    scale invariance of the cosine map, |sim| <= 1, argmax consistency, and agreement with a sampled oracle subset."""
    ops, O = _ops(), _O()
    g = torch.Generator(device="cuda").manual_seed(5)
    p = torch.randn(4096 * 196, 512, device=cuda, generator=g).bfloat16()
    t = torch.nn.functional.normalize(torch.randn(171, 512, device=cuda, generator=g), dim=-1).bfloat16()
    sim, am = ops.patch_text_sim(p, t)
    assert float(sim.abs().max()) <= 1.0 + 1e-3
    assert torch.equal(am.long(), sim.argmax(-1))
    sim2, am2 = ops.patch_text_sim(p * 4, t)                       # power-of-two scale: bf16-exact, cosine unchanged
    assert float((sim2 - sim).abs().max()) < 1e-6 and torch.equal(am2, am)
    idx = torch.randint(0, p.shape[0], (2048,), device=cuda, generator=g)
    ref, _ = O.patch_text_sim(p[idx].cpu()[None], t.cpu())
    assert (sim[idx].cpu() - ref[0]).abs().max().item() < 1e-2


def test_retrieval_golden(cuda):
    ops, O = _ops(), _O()
    gold = np.load(os.path.join(GOLD, "heads_loss.npz"))
    left = torch.tensor(gold["retr_left"]).to(cuda); right = torch.tensor(gold["retr_right"]).to(cuda)
    lg = torch.arange(40, device=cuda); rg = torch.arange(200, device=cuda) // 5
    for prec, tol in ((0, 1e-5), (1, 1e-3)):
        sim = ops.allpairs_sim(left, right, prec)
        assert (sim.cpu() - O.allpairs_sim(left.cpu(), right.cpu())).abs().max().item() < tol
    sim = ops.allpairs_sim(left, right, 0)
    rank = ops.retrieval_rank(sim, lg, rg)
    assert np.array_equal(rank.cpu().numpy(), gold["retr_first"])
    for k, key in ((1, "retr_r1"), (5, "retr_r5"), (10, "retr_r10")):
        assert abs((rank < k).float().mean().item() - float(gold[key])) < 1e-6


def test_retrieval_rank_properties_large(cuda):
    """cfg5-sized similarity (5000 x 25000): rank must equal the count of strictly better items."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(9)
    left = torch.nn.functional.normalize(torch.randn(5000, 512, device=cuda, generator=g), dim=-1)
    right = torch.nn.functional.normalize(torch.randn(25000, 512, device=cuda, generator=g), dim=-1)
    lg = torch.arange(5000, device=cuda); rg = torch.arange(25000, device=cuda) // 5
    sim = ops.allpairs_sim(left, right, 1)
    ref = left @ right.T
    assert (sim - ref).abs().max().item() < 1e-3
    rank = ops.retrieval_rank(sim, lg, rg)
    match = rg[None, :] == lg[:, None]
    best = sim.masked_fill(~match, float("-inf")).max(1)[0]
    assert torch.equal(rank.long(), (sim > best[:, None]).sum(1))


@pytest.mark.parametrize("rows,C", [(12544, 171), (300, 256), (4097, 130), (20000, 20)])
def test_patch_text_sim_pair_equals_single(cuda, rows, C, monkeypatch):
    """The CTA-pair variant (tcgen05 cta_group::2, text matrix split across the two CTAs) and the single-CTA variant of
    the fused kernel compute the same map and the same argmax."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(rows + C)
    p = torch.randn(rows, 512, device=cuda, generator=g).bfloat16()
    t = torch.nn.functional.normalize(torch.randn(C, 512, device=cuda, generator=g), dim=-1).bfloat16()
    monkeypatch.setenv("SIMSEG_PATCH_SIM_CTAS", "1")
    s1, a1 = ops.patch_text_sim(p, t)
    monkeypatch.setenv("SIMSEG_PATCH_SIM_CTAS", "2")
    s2, a2 = ops.patch_text_sim(p, t)
    assert float((s1 - s2).abs().max()) < 1e-6
    assert torch.equal(a2.long(), s2.argmax(-1)) and torch.equal(a1.long(), s1.argmax(-1))


@pytest.mark.parametrize("rows,C", [(12544, 171), (4097, 171), (588, 81), (333, 21), (4099, 150), (50, 255), (7, 1), (2049, 33)])
def test_patch_text_sim_bulk_equals_transposing(cuda, rows, C, monkeypatch):
    """Maps whose rows are not 16-byte aligned (C % 4 != 0) can be written by row-contiguous bulk copies; the result must be
    bit-identical to the transposing epilogue, including ragged last tiles whose byte count is not a multiple of 16
    (those fall back to a plain coalesced copy of the staged rows)."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(rows * 3 + C)
    p = torch.randn(rows, 512, device=cuda, generator=g).bfloat16()
    t = torch.nn.functional.normalize(torch.randn(C, 512, device=cuda, generator=g), dim=-1).bfloat16()
    monkeypatch.setenv("SIMSEG_PATCH_SIM_BULK", "0")
    s0, a0 = ops.patch_text_sim(p, t)
    monkeypatch.setenv("SIMSEG_PATCH_SIM_BULK", "1")               # the heuristic picks it for large many-class maps only
    s1, a1 = ops.patch_text_sim(p, t)
    assert torch.equal(s0, s1) and torch.equal(a0, a1)
    assert torch.equal(a1.long(), s1.argmax(-1))


def test_seg_glue_vs_oracle(cuda):
    """zero-shot class embedding (mean over prompts, renorm), image-level class selection (top-k,
    mean+std threshold, skip ids 0/255, stop below threshold) and the min-max normalised x16 nearest-up-sampled maps —
    GPU kernels against the CPU restatement of tools/seg_evaluation.py:57-75,119-150."""
    ops, O = _ops(), _O()
    g = torch.Generator().manual_seed(9)
    prompts = torch.randn(171, 80, 512, generator=g)
    ce = ops.seg_class_embed(prompts.to(cuda))
    ref_ce = O.zero_shot_class_embedding(prompts)
    assert (ce.cpu() - ref_ce).abs().max().item() < 1e-6
    img = torch.nn.functional.normalize(torch.randn(6, 512, generator=g) + 2.0 * ref_ce[[3, 0, 40, 255 % 171, 100, 7]], dim=-1)
    for topk in (10, 20):
        scores, cand, thr = ops.seg_select(img.to(cuda), ce, topk)
        r_scores, r_cand, r_thr = O.seg_select(img, ref_ce, topk)
        assert (scores.cpu() - r_scores).abs().max().item() < 1e-5
        assert (thr.cpu() - r_thr).abs().max().item() < 1e-5
        assert torch.equal(cand.cpu(), r_cand), (cand.cpu(), r_cand)
    sim = torch.randn(6, 196, 171, generator=g)
    maps = ops.seg_upsample_norm(sim.to(cuda), cand, 14, 14, 16)
    ref_maps = O.seg_norm_maps(sim, r_cand, 14, 14, 16)
    assert maps.shape == (6, 5, 224, 224)
    assert (maps.cpu() - ref_maps).abs().max().item() < 1e-6


def test_segment_pipeline_calls(cuda):
    """simseg_b200.seg.zero_shot_classifier / segment: the tool-level flow end to end on a tiny synthetic problem."""
    from oracle import simseg_oracle as O
    from simseg_b200 import seg
    from simseg_b200.config import load_cfg
    from simseg_b200.pipeline import PIPELINE
    cfg = load_cfg("simseg.vit-s.yaml", ["model.image_encoder.pretrained=False", "model.text_encoder.pretrained=False",
                                          "transforms.input_size=224"])
    model = PIPELINE["clip"](cfg).to(cuda).eval()
    sd = O.make_state_dict(384, 6, seed=0)
    model.load_state_dict(sd)
    gen = torch.Generator().manual_seed(3)
    Cn, P, T = 12, 4, 25
    ids = torch.randint(0, 30522, (Cn, P, T), generator=gen)
    ids[..., 0] = 101
    am = torch.ones(Cn, P, T, dtype=torch.int64)
    am[..., 17:] = 0
    class_emb = seg.zero_shot_classifier(model, ids.to(cuda), am.to(cuda))
    assert class_emb.shape == (Cn, 512)
    assert (class_emb.norm(dim=-1) - 1).abs().max().item() < 1e-5
    # oracle: text tower -> projection/pool/l2norm per caption -> mean over prompts -> renorm
    tt = O.bert_forward(sd, ids.view(-1, T), am.view(-1, T), 12, O.TXT_PREFIX)
    ref = O.zero_shot_class_embedding(O.text_embed(tt, sd["text_projection.linear.weight"], am.view(-1, T)).view(Cn, P, -1))
    assert _cos_rows(class_emb.cpu(), ref) > 0.999
    batch = O.make_batch(2, 25, seed=5)
    sim, amax, scores, cand, maps = seg.segment(model, batch["image"].to(cuda), class_emb, top_cls_num=6)
    assert sim.shape == (2, 196, Cn) and amax.shape == (2, 196) and scores.shape == (2, Cn)
    assert cand.shape == (2, 5) and maps.shape == (2, 5, 224, 224)
    assert torch.equal(amax.long(), sim.argmax(-1))
    live = cand >= 0
    if live.any():
        m = maps[live]
        assert float(m.min()) >= 0.0 and float(m.max()) <= 1.0 + 1e-6


def _cos_rows(a, b):
    a, b = a.double(), b.double()
    return ((a * b).sum(-1) / (a.norm(dim=-1) * b.norm(dim=-1))).min().item()


@pytest.mark.parametrize("g0,g1,D,extra", [(14, 18, 384, 1), (14, 18, 768, 1), (18, 14, 384, 1), (14, 32, 100, 2), (7, 7, 64, 1),
                                           (2, 9, 8, 0)])
def test_pos_embed_bicubic_vs_oracle(cuda, g0, g1, D, extra):
    """utils/interpolate_pe.py:4-27 (checkpoint-load resize, 14x14 -> 18x18 for the 288^2 seg evaluation): the CUDA resize
    against the oracle's F.interpolate(mode='bicubic', align_corners=False) restatement — up-, down-sampling, identity."""
    ops, O = _ops(), _O()
    pe = torch.randn(1, extra + g0 * g0, D, generator=torch.Generator().manual_seed(g0 * 100 + g1))
    out = ops.pos_embed_bicubic(pe.to(cuda), g1, extra)
    ref = O.interpolate_pos_embed(pe, g1 * g1, extra) if g0 != g1 else pe
    assert out.shape == ref.shape
    assert (out.cpu() - ref).abs().max().item() < 1e-5
    assert torch.equal(out[:, :extra].cpu(), pe[:, :extra])


def test_pos_embed_bicubic_vs_reference_fixture(cuda):
    """The same resize against the output of the reference's own interpolate_pos_embed (tests/golden/heads_loss.npz)."""
    z = np.load(os.path.join(GOLD, "heads_loss.npz"))
    out = _ops().pos_embed_bicubic(torch.tensor(z["pe_in"]).to(cuda), 18, 1)
    assert out.shape == (1, 325, 16)
    assert np.abs(out.cpu().numpy() - z["pe_out"]).max() < 2e-5


def test_interpolate_pos_embed_model_288(cuda):
    """seg.interpolate_pos_embed has the reference's call signature; a 224^2 checkpoint loaded into a 288^2 model gives the
    oracle's image tokens (S = 325, the geometry tools/seg_evaluation.py actually evaluates at)."""
    from oracle import simseg_oracle as O
    from simseg_b200 import seg
    from simseg_b200.config import load_cfg
    from simseg_b200.pipeline import PIPELINE
    cfg = load_cfg("simseg.vit-s.yaml", ["model.image_encoder.pretrained=False", "model.text_encoder.pretrained=False",
                                          "transforms.input_size=288"])
    model = PIPELINE["clip"](cfg).to(cuda).eval()
    sd = O.make_state_dict(384, 6, seed=0)                      # a 224^2 checkpoint: pos_embed (1, 197, 384)
    key = "image_encoder.model.model.pos_embed"
    vis = model.image_encoder.model.model
    assert vis.patch_embed.num_patches == 324
    same = seg.interpolate_pos_embed(model.state_dict()[key], vis)
    assert same.shape == (1, 325, 384)                          # grids agree -> returned unchanged
    new_pe = seg.interpolate_pos_embed(sd[key], vis)            # host tensor in, host tensor out
    assert new_pe.shape == (1, 325, 384) and new_pe.device == sd[key].device
    ref_pe = O.interpolate_pos_embed(sd[key], 324)
    assert (new_pe - ref_pe).abs().max().item() < 1e-5
    sd = dict(sd)
    sd[key] = new_pe
    model.load_state_dict(sd)
    image = O.make_batch(2, 25, img_size=288, seed=11)["image"]
    with torch.no_grad():
        feat = model.forward_image_feature(image.to(cuda))
    assert feat.shape == (2, 324, 384)
    sd_ref = dict(sd)
    sd_ref[key] = ref_pe
    ref = O.vit_forward(sd_ref, image, 6, O.IMG_PREFIX)[:, 1:]
    assert _cos_rows(feat.float().cpu().reshape(-1, 384), ref.reshape(-1, 384)) > 0.999
