"""End-to-end parity of the drop-in CLIPModel (bf16 tensor-core path) against the fp32 CPU oracle and the golden
fixture produced by the REFERENCE CLIPModel (tests/golden/clip_vit_s.npz, oracle/make_golden.py).

The product path rounds GEMM operands to bf16 (the reference's autocast contract), the oracle is fp32, so the
end-to-end bars are: embedding cosine to the reference >= 0.9995 and max |diff| <= 4e-3 (components ~0.044),
loss within 2e-2, gradients with cosine >= 0.99 and relative L2 error <= 0.1.  Kernel-level bars (1e-3 fp32 /
1e-2 bf16 on logits and maps, given identical inputs) are in test_heads_loss_gpu.py.
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _build(cuda, yaml="simseg.vit-s.yaml", extra=()):
    from simseg_b200.config import load_cfg
    from simseg_b200.pipeline import PIPELINE
    cfg = load_cfg(yaml, ["model.image_encoder.pretrained=False", "model.text_encoder.pretrained=False",
                          "transforms.input_size=224"] + list(extra))
    return PIPELINE["clip"](cfg).to(cuda), cfg


def _cos(a, b):
    a, b = a.flatten().double(), b.flatten().double()
    return (a @ b / (a.norm() * b.norm() + 1e-30)).item()


def test_state_dict_keys_match_reference_naming(cuda):
    from oracle import simseg_oracle as O
    model, _ = _build(cuda)
    sd = O.make_state_dict(384, 6)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    assert not unexpected and not missing, (missing, unexpected)


def test_clip_vit_s_forward_backward_vs_reference_fixture_and_oracle(cuda):
    from oracle import simseg_oracle as O
    gold = np.load(os.path.join(GOLD, "clip_vit_s.npz"))
    model, _ = _build(cuda)
    sd = O.make_state_dict(384, 6, seed=0)
    model.load_state_dict(sd, strict=True)
    batch = O.make_batch(8, 25, seed=1234)
    gb = {k: v.to(cuda) for k, v in batch.items()}
    # ---- inference API
    with torch.no_grad():
        img_e, txt_e = model(gb, embeddings="all")
        tok = model.forward_image_feature(gb["image"])
    assert tok.shape == (8, 196, 384)
    assert np.abs(tok[:, :4, :32].cpu().numpy() - gold["clip_tokens_head"]).max() < 6e-2
    for got, key in ((img_e, "clip_img_emb"), (txt_e, "clip_txt_emb")):
        ref = torch.tensor(gold[key])
        assert _cos(got.cpu(), ref) > 0.9995
        assert (got.cpu() - ref).abs().max().item() < 4e-3
    # ---- training step
    model.zero_grad(set_to_none=True)
    loss_dict, i2t, t2i = model(gb)
    loss = loss_dict["nce_loss"]
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - float(gold["clip_loss"])) < 2e-2
    assert abs(i2t.item() - float(gold["clip_i2t"])) <= 1 / 8 + 1e-6 and abs(t2i.item() - float(gold["clip_t2i"])) <= 1 / 8 + 1e-6
    gn = float(gold["grad_norm/image_projection.linear.weight"])
    named = dict(model.named_parameters())
    assert abs(named["image_projection.linear.weight"].grad.norm().item() - gn) / gn < 0.1
    assert abs(named["loss.temperature"].grad.item() - float(gold["clip_dtemp"])) / abs(float(gold["clip_dtemp"])) < 0.1


def _grad_table(model, sd, batch, heads):
    from oracle import simseg_oracle as O
    sdg = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    l_o, _, _ = O.clip_train_forward(sdg, batch, heads)
    l_o.backward()
    rows = []
    for k, p in model.named_parameters():
        ref = sdg[k].grad
        assert p.grad is not None, k
        g = p.grad.cpu()
        rows.append((k, _cos(g, ref), ((g - ref).norm() / (ref.norm() + 1e-30)).item(), ref.norm().item()))
    return l_o.item(), rows


def test_tower_and_head_backward_fixed_cotangent_vs_oracle(cuda):
    """Backward kernels in isolation: the same random cotangents are pushed into the (image, text) embeddings of
    ``forward(batch, embeddings='all')`` on both sides, so every one of the 349 tower/head parameter gradients is
    compared without the loss's own conditioning.  Two bars per parameter: (i) relative L2 error to the fp32 oracle
    <= 0.2 with cosine >= 0.98, and (ii) no worse than 1.5x (+0.01) the error the ORACLE ITSELF shows when it is run
    under ``torch.autocast(bfloat16)`` on the GPU — the reference's own precision contract
    (``tasks/clip/clip_runner.py:226-228``).  Measured: mean 0.059 (ours) vs 0.053 (autocast); the weights feeding
    GELU / LayerNorm-scaled GEMMs sit near 0.10-0.15 on both sides (bf16 activations), so (i) alone cannot be tighter."""
    from oracle import simseg_oracle as O
    model, _ = _build(cuda)
    sd = O.make_state_dict(384, 6, seed=0)
    model.load_state_dict(sd, strict=True)
    batch = O.make_batch(8, 25, seed=1234)
    gb = {k: v.to(cuda) for k, v in batch.items()}
    g = torch.Generator().manual_seed(99)
    gi, gt = torch.randn(8, 512, generator=g).to(cuda), torch.randn(8, 512, generator=g).to(cuda)
    model.zero_grad(set_to_none=True)
    ie, te = model(gb, embeddings="all")
    torch.autograd.backward([ie, te], [gi, gt])
    torch.cuda.synchronize()

    def oracle_grads(autocast):
        sdg = {k: v.to(cuda).clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            oi, ot = O.clip_embeddings(sdg, gb, 6)
        torch.autograd.backward([oi.float(), ot.float()], [gi, gt])
        return {k: v.grad for k, v in sdg.items() if v.grad is not None}

    tf32 = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        ref, amp = oracle_grads(False), oracle_grads(True)
    finally:
        torch.backends.cuda.matmul.allow_tf32 = tf32
    rel = lambda a, b: ((a - b).norm() / (b.norm() + 1e-30)).item()
    rows = []
    for k, p in model.named_parameters():
        if k == "loss.temperature" or ref[k].norm().item() <= 1e-7:
            continue
        rows.append((k, _cos(p.grad, ref[k]), rel(p.grad, ref[k]), rel(amp[k], ref[k])))
    bad = sorted(rows, key=lambda r: r[2] - 1.5 * r[3], reverse=True)[:8]
    print("worst gradients (name, cos, rel ours, rel autocast):", *bad, sep="\n  ")
    assert len(rows) > 300
    assert all(r[1] > 0.98 and r[2] < 0.2 for r in rows), bad
    assert all(r[2] <= 1.5 * r[3] + 0.01 for r in rows), bad
    mean_ours, mean_amp = sum(r[2] for r in rows) / len(rows), sum(r[3] for r in rows) / len(rows)
    assert mean_ours <= 1.25 * mean_amp, (mean_ours, mean_amp)


@pytest.mark.parametrize("temperature,min_cos,max_rel", [(0.5, 0.98, 0.2), (0.02, 0.98, 0.2)])
def test_every_parameter_gradient_vs_oracle(cuda, temperature, min_cos, max_rel):
    """All 350 parameter gradients of the full training loss against fp32 autograd through the oracle.  bf16-level
    forward differences (1e-3 on a cosine) are amplified by the loss (x50 logits at the shipped temperature 0.02;
    near-collinear embeddings at any temperature), so the bars here are looser than in the fixed-cotangent test."""
    from oracle import simseg_oracle as O
    model, _ = _build(cuda)
    sd = O.make_state_dict(384, 6, seed=0)
    sd["loss.temperature"] = torch.tensor(temperature)
    model.load_state_dict(sd, strict=True)
    batch = O.make_batch(8, 25, seed=1234)
    gb = {k: v.to(cuda) for k, v in batch.items()}
    model.zero_grad(set_to_none=True)
    loss = model(gb)[0]["nce_loss"]
    loss.backward()
    torch.cuda.synchronize()
    l_o, rows = _grad_table(model, sd, batch, 6)
    assert abs(loss.item() - l_o) < 2e-2
    rows = [r for r in rows if r[3] > 1e-7]
    bad = sorted(rows, key=lambda r: r[1])[:8]
    print("worst gradients (name, cos, rel, |ref|):", *bad, sep="\n  ")
    assert all(r[1] > min_cos and r[2] < max_rel for r in rows), bad


@pytest.mark.parametrize("fused", [False, True])
def test_optimizer_step_changes_outputs_and_cache_refreshes(cuda, fused):
    """bf16 weight copies must follow the fp32 masters after ANY optimizer step.  ``fused=True`` updates parameters
    without bumping ``Tensor._version`` (round 2 found the copies going stale under it), so every step is checked
    against a freshly built model holding the same weights."""
    from oracle import simseg_oracle as O
    model, _ = _build(cuda)
    model.load_state_dict(O.make_state_dict(384, 6, seed=0))
    gb = {k: v.to(cuda) for k, v in O.make_batch(4, 25, seed=5).items()}
    opt = torch.optim.AdamW(model.parameters(), lr=2e-5, fused=fused)
    fresh, _ = _build(cuda)
    l0 = None
    for _ in range(4):
        opt.zero_grad(set_to_none=True)
        loss = model(gb)[0]["nce_loss"]
        loss.backward()
        opt.step()
        l0 = loss.item() if l0 is None else l0
        fresh.load_state_dict(model.state_dict())
        with torch.no_grad():
            a, b = model(gb)[0]["nce_loss"].item(), fresh(gb)[0]["nce_loss"].item()
        assert abs(a - b) < 1e-5, (a, b)
    assert loss.item() < l0            # the same batch gets easier


def test_seg_map_through_reference_tool_calls(cuda):
    """The call sequence of tools/seg_evaluation.py:99-102,111-112,136 against the oracle."""
    from oracle import simseg_oracle as O
    from simseg_b200 import ops
    model, _ = _build(cuda)
    sd = O.make_state_dict(384, 6, seed=0)
    model.load_state_dict(sd)
    batch = O.make_batch(2, 25, seed=77)
    with torch.no_grad():
        feat = model.forward_image_feature(batch["image"].to(cuda))            # (B,196,384)
        pooled = model.forward_image_project(feat)                              # (B,512)
        proj = model.image_projection(feat)                                     # (B,196,512)
    assert proj.shape == (2, 196, 512) and pooled.shape == (2, 512)
    text = torch.nn.functional.normalize(torch.randn(20, 512, generator=torch.Generator().manual_seed(1)), dim=-1)
    sim, am = ops.patch_text_sim(proj.contiguous(), text.to(cuda))
    tok = O.vit_forward(sd, batch["image"], 6, O.IMG_PREFIX)
    ref_sim, _ = O.patch_text_sim(O.simple_projection(tok[:, 1:], sd["image_projection.linear.weight"]), text)
    assert (sim.cpu() - ref_sim).abs().max().item() < 1e-2
    assert torch.equal(am.long(), sim.argmax(-1))


def test_clip_vit_b_forward_backward_and_seg_map_vs_oracle(cuda):
    """BASELINE configs[2]/[3] model (ViT-B/16 + BERT-base) at a batch the CPU oracle finishes in seconds: embeddings, loss,
    a sample of parameter gradients, the 171-class patch-text map (CTA-pair kernel) and its argmax mask."""
    from oracle import simseg_oracle as O
    from simseg_b200 import ops
    model, _ = _build(cuda, yaml="simseg.vit-b.yaml")
    sd = O.make_state_dict(768, 12, seed=3)
    model.load_state_dict(sd, strict=True)
    batch = O.make_batch(4, 25, seed=4321)
    gb = {k: v.to(cuda) for k, v in batch.items()}
    with torch.no_grad():
        img_e, txt_e = model(gb, embeddings="all")
        feat = model.forward_image_feature(gb["image"])
        proj = model.image_projection(feat)
    tok = O.vit_forward(sd, batch["image"], 12, O.IMG_PREFIX)
    assert feat.shape == (4, 196, 768)
    assert (feat.cpu() - tok[:, 1:]).abs().max().item() < 0.15              # bf16 operands through 12 blocks, values ~ +-4
    ref_proj = O.simple_projection(tok[:, 1:], sd["image_projection.linear.weight"])
    text = torch.nn.functional.normalize(torch.randn(171, 512, generator=torch.Generator().manual_seed(5)), dim=-1)
    sim, am = ops.patch_text_sim(proj.contiguous(), text.to(cuda))
    ref_sim, _ = O.patch_text_sim(ref_proj, text)
    assert (sim.cpu() - ref_sim).abs().max().item() < 1e-2
    assert torch.equal(am.long(), sim.argmax(-1))
    top2 = ref_sim.topk(2, -1)[0]
    safe = (top2[..., 0] - top2[..., 1]) > 2e-2                                # mask is bit-exact wherever the oracle margin allows
    assert torch.equal(am.cpu().long()[safe], ref_sim.argmax(-1)[safe])
    # training step
    model.zero_grad(set_to_none=True)
    loss = model(gb)[0]["nce_loss"]
    loss.backward()
    torch.cuda.synchronize()
    l_o, rows = _grad_table(model, sd, batch, 12)
    assert abs(loss.item() - l_o) < 2e-2
    big = [r for r in rows if r[3] > 1e-4]
    assert len(big) > 100
    cos = sorted(r[1] for r in big)
    assert cos[len(cos) // 20] > 0.97, cos[:10]                                 # 95 % of the parameter gradients
    ref_i, ref_t = O.clip_embeddings(sd, batch, 12) if hasattr(O, "clip_embeddings") else (None, None)
    if ref_i is not None:
        assert _cos(img_e.cpu(), ref_i) > 0.999 and _cos(txt_e.cpu(), ref_t) > 0.999


def test_clip_vit_s_77_token_captions_vs_oracle(cuda):
    """BASELINE configs[0] geometry at model level: 77-token captions (ragged lengths 8..77), ViT-S/16, 20-class map —
    embeddings, image-text logits (1e-2 bf16 bar on cosines x 1/0.02 is the loss bar 2e-2 used elsewhere), loss and a
    sample of BERT gradients (the T = 77 attention path packs fewer captions per tile than T = 25)."""
    from oracle import simseg_oracle as O
    from simseg_b200 import ops
    model, _ = _build(cuda)
    sd = O.make_state_dict(384, 6, seed=0)
    model.load_state_dict(sd, strict=True)
    batch = O.make_batch(6, 77, seed=2024)
    batch["attention_mask"][0] = 1                                            # one full-length caption
    gb = {k: v.to(cuda) for k, v in batch.items()}
    with torch.no_grad():
        img_e, txt_e = model(gb, embeddings="all")
        tfeat = model.forward_text_feature(gb["input_ids"], gb["attention_mask"])
        proj = model.image_projection(model.forward_image_feature(gb["image"]))
    ref_i, ref_t = O.clip_embeddings(sd, batch, 6)
    ref_tok = O.bert_forward(sd, batch["input_ids"], batch["attention_mask"], 12, O.TXT_PREFIX)
    assert tfeat.shape == (6, 77, 768)
    valid = batch["attention_mask"].bool()
    assert (tfeat.cpu()[valid] - ref_tok[valid]).abs().max().item() < 0.1     # LayerNorm-ed hidden states, values ~ +-3
    assert _cos(img_e.cpu(), ref_i) > 0.9995 and _cos(txt_e.cpu(), ref_t) > 0.9995
    assert (txt_e.cpu() - ref_t).abs().max().item() < 4e-3
    cos_ours, cos_ref = img_e.cpu() @ txt_e.cpu().T, ref_i @ ref_t.T
    assert (cos_ours - cos_ref).abs().max().item() < 1e-2
    text = torch.nn.functional.normalize(torch.randn(20, 512, generator=torch.Generator().manual_seed(1)), dim=-1)
    sim, am = ops.patch_text_sim(proj.contiguous(), text.to(cuda))
    tok = O.vit_forward(sd, batch["image"], 6, O.IMG_PREFIX)
    ref_sim, _ = O.patch_text_sim(O.simple_projection(tok[:, 1:], sd["image_projection.linear.weight"]), text)
    assert (sim.cpu() - ref_sim).abs().max().item() < 1e-2
    model.zero_grad(set_to_none=True)
    loss = model(gb)[0]["nce_loss"]
    loss.backward()
    torch.cuda.synchronize()
    l_o, rows = _grad_table(model, sd, batch, 6)
    assert abs(loss.item() - l_o) < 2e-2
    txt_rows = [r for r in rows if r[0].startswith("text_") and r[3] > 1e-6]
    assert len(txt_rows) > 150
    cs = sorted(r[1] for r in txt_rows)
    assert cs[len(cs) // 20] > 0.97, cs[:8]
