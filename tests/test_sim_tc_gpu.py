"""The tensor-core similarity family (round 2): InfoNCE and retrieval ranking on the tcgen05 GEMM engine with split-bf16
("fp32-grade") products and the score matrix consumed inside the GEMM epilogue.

Parity bars: logits within 1e-3 of the fp32 reference at the shipped temperature 0.02 (north star), loss 1e-4 on the reference
fixtures, gradients 1e-4 relative, first-max accuracy and retrieval ranks bit-exact on the reference fixtures; at full size
(cfg2 4096 x 4096, cfg3 1024 x 8192, cfg5 5000 x 25000) size-independent properties + an fp64 torch restatement on the GPU.
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


def _ref_nce(f1, f2g, temp, off):
    """fp64 restatement of mml_loss.py:56,73-77 on the GPU (+ gradients by autograd)."""
    f1 = f1.double().requires_grad_(True)
    f2g = f2g.double().requires_grad_(True)
    t = temp.double().clone().requires_grad_(True)
    logits = f1 @ f2g.T / torch.clamp(t, 0.001, 0.5)
    tgt = torch.arange(off, off + f1.shape[0], device=f1.device)
    rows = torch.nn.functional.cross_entropy(logits, tgt, reduction="none")
    return logits, rows, tgt, f1, f2g, t


def test_allpairs_split_products_reach_fp32_grade(cuda):
    """The arithmetic itself: split-bf16 cosines vs fp64 on correlated unit-norm embeddings (the case single-pass tf32
    misses by 50x, see test_heads_loss_gpu.py)."""
    ops = _ops()
    gold = np.load(os.path.join(GOLD, "heads_loss.npz"))
    img = torch.tensor(gold["heads_img_emb"]).to(cuda); txt = torch.tensor(gold["heads_txt_emb"]).to(cuda)
    cos = ops.allpairs_sim_split(img, txt)
    ref = (img.double() @ txt.double().T)
    assert (cos.double() - ref).abs().max().item() * 50 < 1e-3          # logits = cos / 0.02
    g = torch.Generator(device="cuda").manual_seed(1)
    a = torch.nn.functional.normalize(torch.randn(1000, 512, device=cuda, generator=g) + 3.0, dim=-1)   # cos ~ 0.9: no cancellation
    b = torch.nn.functional.normalize(torch.randn(777, 512, device=cuda, generator=g) + 3.0, dim=-1)
    err = (ops.allpairs_sim_split(a, b).double() - a.double() @ b.double().T).abs().max().item()
    print("split-bf16 max |cos err|", err)
    assert err < 2e-5 and err * 50 < 1e-3


def test_infonce_fused_vs_golden_and_oracle(cuda):
    ops, O = _ops(), _O()
    gold = np.load(os.path.join(GOLD, "heads_loss.npz"))
    img = torch.tensor(gold["heads_img_emb"]); txt = torch.tensor(gold["heads_txt_emb"])
    temp = torch.tensor(0.02)
    b = img.shape[0]
    gi, gt, gtemp = img.to(cuda), txt.to(cuda), temp.to(cuda)
    l1, lse1, am1, ws1 = ops.infonce_fused_fwd(gi, gt, gtemp, 0)
    l2, lse2, am2, ws2 = ops.infonce_fused_fwd(gt, gi, gtemp, 0)
    loss = 0.5 * (l1.mean() + l2.mean())
    _, _, ref_logits, ref_rows = O.nce_direction(img, txt, temp, 0)
    assert (l1.cpu() - ref_rows).abs().max().item() < 1e-3
    assert (lse1.cpu() - torch.logsumexp(ref_logits, 1)).abs().max().item() < 1e-3
    assert abs(loss.item() - float(gold["nce_loss"])) < 1e-4
    assert torch.equal(am1.cpu().long(), ref_logits.argmax(1))
    assert abs((am1.cpu() == torch.arange(b)).float().mean().item() - float(gold["nce_i2t"])) < 1e-6
    assert abs((am2.cpu() == torch.arange(b)).float().mean().item() - float(gold["nce_t2i"])) < 1e-6
    dimg_g = torch.zeros_like(gi); dtxt_g = torch.zeros_like(gt); dtemp = torch.zeros((), device=cuda)
    d_img = ops.infonce_fused_bwd(b, b, 512, gtemp, 0, lse1, 0.5 / b, ws1, dtxt_g, dtemp)
    d_txt = ops.infonce_fused_bwd(b, b, 512, gtemp, 0, lse2, 0.5 / b, ws2, dimg_g, dtemp)
    for got, ref in ((d_img + dimg_g, gold["nce_dimg"]), (d_txt + dtxt_g, gold["nce_dtxt"])):
        ref = torch.tensor(ref)
        assert ((got.cpu() - ref).abs().max() / ref.abs().max()).item() < 1e-4
    assert abs(dtemp.item() - float(gold["nce_dtemp"])) / abs(float(gold["nce_dtemp"])) < 1e-3


def test_infonce_fused_global_reduce_two_rank_fixture(cuda):
    """Row offsets (targets = b*rank + i, mml_loss.py:75) and the accumulated gathered-operand gradient against the
    reference's own 2-rank gloo run (tests/golden/global_reduce.npz)."""
    ops = _ops()
    gold = np.load(os.path.join(GOLD, "global_reduce.npz"))
    img = torch.tensor(gold["gr_img"]).to(cuda); txt = torch.tensor(gold["gr_txt"]).to(cuda)
    temp = torch.tensor(0.02, device=cuda)
    W = 2
    b = img.shape[0] // W
    dimg_g = torch.zeros_like(img); dtxt_g = torch.zeros_like(txt)
    dloc_i, dloc_t = [], []
    for r in range(W):
        li, lt = img[r * b:(r + 1) * b].contiguous(), txt[r * b:(r + 1) * b].contiguous()
        l1, lse1, _, ws1 = ops.infonce_fused_fwd(li, txt, temp, r * b)
        l2, lse2, _, ws2 = ops.infonce_fused_fwd(lt, img, temp, r * b)
        loss = 0.5 * (l1.mean() + l2.mean())
        assert abs(loss.item() - float(gold[f"gr_loss_{r}"])) < 1e-4
        dloc_i.append(ops.infonce_fused_bwd(b, W * b, 512, temp, r * b, lse1, 0.5 / b, ws1, dtxt_g, None))
        dloc_t.append(ops.infonce_fused_bwd(b, W * b, 512, temp, r * b, lse2, 0.5 / b, ws2, dimg_g, None))
    for r in range(W):
        gi = dloc_i[r] + dimg_g[r * b:(r + 1) * b]
        gt = dloc_t[r] + dtxt_g[r * b:(r + 1) * b]
        ri, rt = torch.tensor(gold[f"gr_dimg_{r}"]), torch.tensor(gold[f"gr_dtxt_{r}"])
        assert ((gi.cpu() - ri).abs().max() / ri.abs().max()).item() < 1e-4
        assert ((gt.cpu() - rt).abs().max() / rt.abs().max()).item() < 1e-4


@pytest.mark.parametrize("b,Bg,off,temp", [(1, 1, 0, 0.02), (37, 111, 40, 0.02), (130, 257, 127, 0.07), (300, 1200, 600, 0.02),
                                            (129, 129, 0, 0.0005), (64, 2049, 1985, 0.9)])
def test_infonce_fused_ragged_shapes_vs_fp64(cuda, b, Bg, off, temp):
    """Ragged tile edges (rows / columns not multiples of 128 / 256 / 8), row offsets, and temperatures outside the clamp
    range [0.001, 0.5] (clamped at use, zero temperature gradient — mml_loss.py:56)."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(b * 1000 + Bg)
    base = torch.randn(1, 512, device=cuda, generator=g)
    f2g = torch.nn.functional.normalize(torch.randn(Bg, 512, device=cuda, generator=g) + base, dim=-1)
    f1 = torch.nn.functional.normalize(f2g[off:off + b] + 0.3 * torch.randn(b, 512, device=cuda, generator=g), dim=-1).contiguous()
    t = torch.tensor(temp, device=cuda)
    rows, lse, am, ws = ops.infonce_fused_fwd(f1, f2g, t, off)
    logits, ref_rows, tgt, r1, r2, rt = _ref_nce(f1, f2g, t, off)
    tol = 1e-3 * max(1.0, 0.02 / min(max(temp, 0.001), 0.5))              # the 1e-3 bar is stated at temperature 0.02
    assert (rows.double() - ref_rows).abs().max().item() < tol
    assert (lse.double() - torch.logsumexp(logits, 1)).abs().max().item() < tol
    top2 = logits.topk(min(2, Bg), 1)[0]
    safe = (top2[:, 0] - top2[:, -1]) > 2 * tol if Bg > 1 else torch.ones(b, dtype=torch.bool, device=cuda)
    assert torch.equal(am.long()[safe], logits.argmax(1)[safe])
    ref_rows.mean().backward()
    df2g = torch.zeros_like(f2g); dtemp = torch.zeros((), device=cuda)
    df1 = ops.infonce_fused_bwd(b, Bg, 512, t, off, lse, 1.0 / b, ws, df2g, dtemp)
    scale = 1.0 / (b * min(max(temp, 0.001), 0.5))                          # |dL/dcos| <= 1/(b t): the natural gradient scale
    for got, ref in ((df1, r1.grad), (df2g, r2.grad)):                     # (b = Bg = 1: the true gradient is exactly 0)
        assert ((got.double() - ref).abs().max() / max(ref.abs().max().item(), 0.05 * scale)).item() < 2e-4
    if 0.001 <= temp <= 0.5:
        tc = min(max(temp, 0.001), 0.5)
        assert abs(dtemp.item() - rt.grad.item()) / max(abs(rt.grad.item()), 0.05 / (tc * tc)) < 1e-3
    else:
        assert dtemp.item() == 0.0 and rt.grad.item() == 0.0


@pytest.mark.parametrize("b,Bg,off", [(4096, 4096, 0), (1024, 8192, 3072)])
def test_infonce_fused_full_size_properties(cuda, b, Bg, off):
    """BASELINE configs[1] (one GPU: 4096 x 4096) and configs[2] (per rank: 1024 rows x 8192 gathered columns, rank 3)."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(Bg)
    f2g = torch.nn.functional.normalize(torch.randn(Bg, 512, device=cuda, generator=g), dim=-1)
    # cos(f1_i, target) ~ 0.2 -> target logit ~ 10 against 4096 / 8192 distractors of std 2.2: a mid-training loss (~1),
    # so the softmax is neither uniform nor saturated
    f1 = torch.nn.functional.normalize(0.2 * f2g[off:off + b] + torch.nn.functional.normalize(
        torch.randn(b, 512, device=cuda, generator=g), dim=-1), dim=-1).contiguous()
    t = torch.tensor(0.02, device=cuda)
    rows, lse, am, ws = ops.infonce_fused_fwd(f1, f2g, t, off)
    logits, ref_rows, tgt, r1, r2, rt = _ref_nce(f1, f2g, t, off)
    assert (rows.double() - ref_rows).abs().max().item() < 1e-3
    assert 0.05 < ref_rows.mean().item() < 6.0
    top2 = logits.topk(2, 1)[0]
    safe = (top2[:, 0] - top2[:, 1]) > 2e-3
    assert safe.float().mean().item() > 0.99 and torch.equal(am.long()[safe], logits.argmax(1)[safe])
    ref_rows.mean().backward()
    df2g = torch.zeros_like(f2g); dtemp = torch.zeros((), device=cuda)
    df1 = ops.infonce_fused_bwd(b, Bg, 512, t, off, lse, 1.0 / b, ws, df2g, dtemp)
    assert ((df1.double() - r1.grad).abs().max() / r1.grad.abs().max()).item() < 2e-4
    assert ((df2g.double() - r2.grad).abs().max() / r2.grad.abs().max()).item() < 2e-4
    assert abs(dtemp.item() - rt.grad.item()) / abs(rt.grad.item()) < 1e-3
    # size-independent property: the loss is invariant to scaling both operands by 2 and the temperature by 4
    # (power-of-two scales: the bf16 hi/lo splits and every product are bit-identical up to the exponent)
    rows2, _, _, _ = ops.infonce_fused_fwd((2 * f1).contiguous(), (2 * f2g).contiguous(), torch.tensor(0.08, device=cuda), off)
    assert (rows2 - rows).abs().max().item() < 1e-4


def test_model_loss_uses_tensor_core_infonce(cuda):
    """``NCE.forward`` (default precision) == the exact-fp32 SIMT cross-check path, forward and gradients."""
    from simseg_b200._lib import PREC_FP32, PREC_SPLIT_BF16
    from simseg_b200.config import load_cfg
    from simseg_b200.pipeline import NCE
    cfg = load_cfg("simseg.vit-s.yaml", [])
    g = torch.Generator(device="cuda").manual_seed(3)
    img = torch.nn.functional.normalize(torch.randn(96, 512, device=cuda, generator=g), dim=-1)
    txt = torch.nn.functional.normalize(img + 0.5 * torch.randn(96, 512, device=cuda, generator=g), dim=-1)
    out = {}
    for prec in (PREC_SPLIT_BF16, PREC_FP32):
        loss_mod = NCE(cfg, 0).to(cuda)
        assert loss_mod.precision == PREC_SPLIT_BF16
        loss_mod.precision = prec
        i, t = img.clone().requires_grad_(True), txt.clone().requires_grad_(True)
        l1, a1 = loss_mod(i, t)
        l2, a2 = loss_mod(t, i)
        (0.5 * (l1 + l2)).backward()
        out[prec] = (l1.item(), l2.item(), a1.item(), a2.item(), i.grad.clone(), t.grad.clone(), loss_mod.temperature.grad.item())
    a, b = out[PREC_SPLIT_BF16], out[PREC_FP32]
    assert abs(a[0] - b[0]) < 1e-4 and abs(a[1] - b[1]) < 1e-4 and a[2] == b[2] and a[3] == b[3]
    assert ((a[4] - b[4]).abs().max() / b[4].abs().max()).item() < 1e-4
    assert ((a[5] - b[5]).abs().max() / b[5].abs().max()).item() < 1e-4
    assert abs(a[6] - b[6]) / abs(b[6]) < 1e-3


# ------------------------------------------------------------------------------------------------ retrieval
def test_retrieval_fused_golden(cuda):
    """Reference fixture (EmbANN + RetrievalMetric run on the reference's own code): ranks bit-exact, R@K equal."""
    from simseg_b200.retrieval import IndexedEmbInfo, RetrievalMetric
    ops = _ops()
    gold = np.load(os.path.join(GOLD, "heads_loss.npz"))
    left = torch.tensor(gold["retr_left"]).to(cuda); right = torch.tensor(gold["retr_right"]).to(cuda)
    lg = torch.arange(40, device=cuda); rg = torch.arange(200, device=cuda) // 5
    rank = ops.retrieval_rank_fused(left, right, lg, rg)
    assert np.array_equal(rank.cpu().numpy(), gold["retr_first"])
    res = RetrievalMetric(with_prefix=False)(IndexedEmbInfo("image", lg, left), IndexedEmbInfo("text", rg, right))
    for k in (1, 5, 10):
        assert abs(res[f"R@{k}"] - float(gold[f"retr_r{k}"])) < 1e-6
    res = RetrievalMetric()(IndexedEmbInfo("image", lg, left), IndexedEmbInfo("text", rg, right))
    assert "[image] to [text]: R@1" in res


def _rank_ref(sim, lg, rg):
    """Definition (stable descending order) on a materialised score matrix, any arithmetic."""
    match = rg[None, :] == lg[:, None]
    masked = sim.masked_fill(~match, float("-inf"))
    best, _ = masked.max(1)
    cols = torch.arange(sim.shape[1], device=sim.device)[None, :]
    first = torch.where(masked == best[:, None], cols, sim.shape[1]).min(1)[0]       # lowest column among the best matches
    beats = ((sim > best[:, None]) | ((sim == best[:, None]) & (cols < first[:, None]))) & ~match
    rank = beats.sum(1)
    return torch.where(match.any(1), rank, torch.full_like(rank, -1))


@pytest.mark.parametrize("M,Nr,per", [(5000, 25000, 5), (300, 777, 3), (129, 4000, 1), (1, 1, 1), (2049, 1000, 7)])
def test_retrieval_fused_equals_definition_on_its_own_scores(cuda, M, Nr, per):
    """BASELINE configs[4] size and ragged shapes: the fused ranks must equal the definition evaluated on the SAME
    split-product scores (materialising variant ``allpairs_sim_split`` — identical arithmetic, so bit-exact), and differ from
    the definition on exact-fp32 scores only where the fp32 margin is below the split-product error."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(M + Nr)
    left = torch.nn.functional.normalize(torch.randn(M, 512, device=cuda, generator=g), dim=-1)
    right = torch.nn.functional.normalize(torch.randn(Nr, 512, device=cuda, generator=g), dim=-1)
    lg = torch.arange(M, device=cuda)
    rg = torch.arange(Nr, device=cuda) // per            # rows beyond Nr/per have no match -> -1
    rank = ops.retrieval_rank_fused(left, right, lg, rg)
    sim = ops.allpairs_sim_split(left, right)
    assert torch.equal(rank.long(), _rank_ref(sim, lg, rg))
    exact = _rank_ref(ops.allpairs_sim(left, right, 0), lg, rg)
    # random data puts the best match in the dense bulk of the score distribution (~2e5 scores per unit of cosine), where
    # ~3 % of the rows have another score within the 1e-6 split-product error of s*; real retrieval has s* in the sparse tail
    diff = (rank.long() != exact)
    print("rows whose rank differs from the exact-fp32 definition:", diff.float().mean().item())
    assert diff.float().mean().item() < 0.06
    assert (rank.long() - exact).abs().max().item() <= 2
    assert torch.equal(rank < 0, exact < 0)


def test_retrieval_fused_ties_duplicates_and_unsorted_ids(cuda):
    """Collisions as the domain has them: duplicate captions (identical rows -> exactly equal scores -> the lower column
    wins), several matching items per row, unsorted / repeated group ids (tile list falls back to every tile), rows without
    any match."""
    ops = _ops()
    g = torch.Generator(device="cuda").manual_seed(77)
    M, Nr = 513, 3000
    left = torch.nn.functional.normalize(torch.randn(M, 512, device=cuda, generator=g), dim=-1)
    right = torch.nn.functional.normalize(torch.randn(Nr, 512, device=cuda, generator=g), dim=-1)
    lg = torch.randperm(M, device=cuda, generator=g)
    rg = torch.randint(0, M + 50, (Nr,), device=cuda, generator=g)
    # duplicates of matching items placed before and after them, under another group id
    for i in range(0, 200, 7):
        cols = (rg == lg[i]).nonzero().flatten()
        if cols.numel() == 0:
            continue
        j = int(cols[0])
        for tgt in (max(j - 3, 0), min(j + 5, Nr - 1)):
            if rg[tgt] != lg[i]:
                right[tgt] = right[j]
    rank = ops.retrieval_rank_fused(left, right, lg, rg)
    sim = ops.allpairs_sim_split(left, right)
    ref = _rank_ref(sim, lg, rg)
    assert torch.equal(rank.long(), ref)
    assert (rank < 0).any() and (rank >= 0).any()


def test_indexed_emb_info_unique(cuda):
    """``IndexedEmbInfo.unique`` (utils.py:14-19): one row per id, ids ascending, the last duplicate kept."""
    from simseg_b200.retrieval import IndexedEmbInfo
    gid = torch.tensor([5, 3, 5, 9, 3, 3, 1], device=cuda)
    emb = torch.arange(7, device=cuda, dtype=torch.float32)[:, None].repeat(1, 4)
    u = IndexedEmbInfo("image", gid, emb).unique()
    assert u.group_idx.tolist() == [1, 3, 5, 9]
    assert u.emb_mat[:, 0].tolist() == [6.0, 5.0, 2.0, 3.0]
    chunks = list(IndexedEmbInfo("image", gid, emb).to_chunks(3))
    assert [c.emb_mat.shape[0] for c in chunks] == [3, 3, 1]
