"""CPU: the oracle (oracle/simseg_oracle.py) against the committed golden fixtures.

The fixtures under tests/golden/ hold outputs of the REFERENCE's own modules (``oracle/make_golden.py`` imported
them from /root/reference in the build container); nothing here reads /root/reference.  These tests keep the
oracle pinned so that the ``-m gpu`` parity tests compare the CUDA path with a checked checker.
"""
import os

import numpy as np
import torch

from oracle import simseg_oracle as O

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _max(a, b):
    return (torch.as_tensor(a).double() - torch.as_tensor(b).double()).abs().max().item()


def _heads_inputs():
    """Regenerate the seeded inputs of oracle/make_golden.py:heads_and_loss (generator seed 7, same draw order)."""
    g = torch.Generator().manual_seed(7)
    x_img = torch.randn(6, 196, 384, generator=g)
    x_txt = torch.randn(6, 25, 768, generator=g)
    return x_img, x_txt


def test_heads_vs_reference_fixture():
    z = np.load(os.path.join(GOLD, "heads_loss.npz"))
    x_img, x_txt = _heads_inputs()
    assert _max(x_img[:2, :, :64], z["heads_x_img"]) == 0.0           # the regenerated inputs ARE the fixture's
    wi, wt, mask = torch.tensor(z["heads_wi"]), torch.tensor(z["heads_wt"]), torch.tensor(z["heads_mask"])
    img = O.image_embed(torch.cat([x_img[:, :1], x_img], 1), wi, 5)   # drop-CLS + project + top-5 mean + L2norm
    txt = O.text_embed(x_txt, wt, mask, 1)
    assert _max(img, z["heads_img_emb"]) < 2e-5
    assert _max(txt, z["heads_txt_emb"]) < 2e-5


def test_topk_pooling_edge_cases():
    g = torch.Generator().manual_seed(1)
    x = torch.randn(3, 7, 16, generator=g)
    # k shrinks to the shortest caption (pooling.py:61-63); a length-1 caption makes every k behave as k=1
    mask = torch.tensor([[1] * 7, [1] + [0] * 6, [1, 1, 1, 0, 0, 0, 0]])
    got = O.topk_pooling(x, 3, mask)
    want = torch.stack([x[0].max(0)[0], x[1, 0], x[2, :3].max(0)[0]])
    assert _max(got, want) == 0.0
    # k == ntok -> plain mean
    assert _max(O.topk_pooling(x, 7), x.mean(1)) < 1e-6
    # zero vector: L2norm adds eps after the sqrt (normalization.py:9-10) -> 0, not NaN
    assert torch.equal(O.l2norm(torch.zeros(2, 8)), torch.zeros(2, 8))


def test_nce_vs_reference_fixture():
    z = np.load(os.path.join(GOLD, "heads_loss.npz"))
    a = torch.tensor(z["heads_img_emb"]).requires_grad_(True)
    b = torch.tensor(z["heads_txt_emb"]).requires_grad_(True)
    t = torch.tensor(0.02, requires_grad=True)
    loss, i2t, t2i = O.clip_loss(a, b, a, b, t, 0)
    loss.backward()
    assert _max(loss, z["nce_loss"]) < 2e-5
    assert _max(i2t, z["nce_i2t"]) == 0 and _max(t2i, z["nce_t2i"]) == 0
    assert _max(a.grad, z["nce_dimg"]) < 2e-5 and _max(b.grad, z["nce_dtxt"]) < 2e-5
    assert _max(t.grad, z["nce_dtemp"]) < 1e-3


def test_nce_temperature_clamp():
    g = torch.Generator().manual_seed(2)
    a, b = O.l2norm(torch.randn(5, 32, generator=g)), O.l2norm(torch.randn(5, 32, generator=g))
    for raw, eff in ((1e-5, 0.001), (3.0, 0.5)):                      # mml_loss.py:56
        l1, _, lg1, _ = O.nce_direction(a, b, torch.tensor(raw))
        l2, _, lg2, _ = O.nce_direction(a, b, torch.tensor(eff))
        assert _max(l1, l2) == 0 and _max(lg1, lg2) == 0


def test_global_reduce_two_rank_fixture():
    """Global-reduce branch (GatherLayer + targets = arange(b*rank, b*(rank+1))) against the reference run on 2 gloo ranks."""
    z = np.load(os.path.join(GOLD, "global_reduce.npz"))
    img, txt = torch.tensor(z["gr_img"]), torch.tensor(z["gr_txt"])
    b = 6
    for rank in range(2):
        ig, tg = img.clone().requires_grad_(True), txt.clone().requires_grad_(True)
        total = 0
        for r in range(2):
            l, _, _ = O.clip_loss(ig[r * b:(r + 1) * b], tg[r * b:(r + 1) * b], ig, tg, torch.tensor(0.02), r)
            if r == rank:
                assert _max(l, z[f"gr_loss_{rank}"]) < 1e-5
            total = total + l
        total.backward()
        assert _max(ig.grad[rank * b:(rank + 1) * b], z[f"gr_dimg_{rank}"]) < 1e-5
        assert _max(tg.grad[rank * b:(rank + 1) * b], z[f"gr_dtxt_{rank}"]) < 1e-5


def test_retrieval_vs_reference_fixture():
    z = np.load(os.path.join(GOLD, "heads_loss.npz"))
    left, right = torch.tensor(z["retr_left"]), torch.tensor(z["retr_right"])
    lg, rg = torch.arange(40), torch.arange(200) // 5
    has, first = O.retrieval_first_match_rank(O.allpairs_sim(left, right), lg, rg)
    assert bool(has.all()) and torch.equal(first, torch.tensor(z["retr_first"]))
    rec = O.recall_at(has, first)
    for k, key in (("R@1", "retr_r1"), ("R@5", "retr_r5"), ("R@10", "retr_r10")):
        assert abs(rec[k] - float(z[key])) < 1e-7


def test_seg_map_vs_reference_fixture():
    z = np.load(os.path.join(GOLD, "heads_loss.npz"))
    sim, am = O.patch_text_sim(torch.tensor(z["seg_patch"]), torch.tensor(z["seg_text"]))
    assert _max(sim, z["seg_sim"]) < 2e-6
    assert torch.equal(am, torch.tensor(z["seg_argmax"]))
    # an all-zero patch row: F.normalize clamps the norm at 1e-12 -> similarity 0 everywhere, argmax = class 0
    p = torch.tensor(z["seg_patch"]).clone()
    p[0, 3] = 0
    s2, a2 = O.patch_text_sim(p, torch.tensor(z["seg_text"]))
    assert float(s2[0, 3].abs().max()) == 0.0 and int(a2[0, 3]) == 0


def test_pos_embed_interpolation_vs_reference_fixture():
    z = np.load(os.path.join(GOLD, "heads_loss.npz"))
    out = O.interpolate_pos_embed(torch.tensor(z["pe_in"]), 324)      # bicubic is per channel: a 16-channel slice suffices
    assert out.shape == (1, 325, 16)
    assert _max(out, z["pe_out"]) < 2e-5
    assert O.interpolate_pos_embed(torch.tensor(z["pe_in"]), 196) is not None


def test_full_clip_vit_s_vs_reference_fixture():
    """Whole ViT-S + BERT-base + heads + NCE forward/backward (B = 8) against the reference CLIPModel's outputs."""
    z = np.load(os.path.join(GOLD, "clip_vit_s.npz"))
    torch.set_num_threads(os.cpu_count())
    sd = O.make_state_dict(384, 6, seed=0)
    batch = O.make_batch(8, 25, seed=1234)
    sdg = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    loss, i2t, t2i = O.clip_train_forward(sdg, batch, 6)
    loss.backward()
    assert _max(loss, z["clip_loss"]) < 1e-4
    assert _max(i2t, z["clip_i2t"]) == 0 and _max(t2i, z["clip_t2i"]) == 0
    with torch.no_grad():
        img, txt = O.clip_embeddings(sd, batch, 6)
        tok = O.vit_forward(sd, batch["image"], 6, O.IMG_PREFIX)[:, 1:]
    assert _max(img, z["clip_img_emb"]) < 1e-4 and _max(txt, z["clip_txt_emb"]) < 1e-4
    assert _max(tok[:, :4, :32], z["clip_tokens_head"]) < 2e-4
    assert _max(sdg["image_projection.linear.weight"].grad[:8, :32], z["clip_dWi_head"]) < 2e-4 * max(
        1.0, float(np.abs(z["clip_dWi_head"]).max()))
    assert abs(float(sdg["loss.temperature"].grad) - float(z["clip_dtemp"])) < 2e-3 * max(1.0, abs(float(z["clip_dtemp"])))
    for key in z.files:
        if key.startswith("grad_norm/"):
            g = sdg[key[len("grad_norm/"):]].grad
            assert abs(g.norm().item() - float(z[key])) <= 2e-3 * float(z[key]) + 1e-9, key


def test_seg_glue_known_answers():
    """Hand-worked cases of the rules tools/seg_evaluation.py applies (lines cited in the oracle; the outputs of the script's
    own lines are checked in test_seg_glue_vs_reference_fixture): class-embedding mean + renormalisation (:71-72), image-level scores,
    top-k, threshold = mean + unbiased std (:119-124), the scan over the FIRST five of the top-k that skips ids 0 and 255
    without replacing them and stops at the first score below the threshold (:131-147), nearest up-sampling and min-max
    normalisation (:136-147)."""
    # class embedding
    pr = np.arange(24, dtype=np.float32).reshape(2, 3, 4) - 7.0
    m = pr.mean(1)
    want = m / np.linalg.norm(m, axis=-1, keepdims=True)
    assert _max(O.zero_shot_class_embedding(torch.tensor(pr)), want) < 1e-6
    # class selection: identity text matrix => scores = the image embedding itself
    Cn = 300
    text = torch.eye(Cn)
    img = torch.linspace(-0.3, -0.2, Cn).repeat(2, 1)              # distinct, small background scores
    for c, v in ((0, 0.9), (7, 0.8), (255, 0.7), (3, 0.6), (9, 0.5)):
        img[:, c] = v
    # (a) top-10: threshold lands between 0.8 and 0.6 -> class 0 skipped, 7 kept, 255 skipped, 3 below threshold: stop
    scores, cand, thr = O.seg_select(img, text, 10)
    top = np.sort(img[0].numpy())[::-1][:10]
    assert abs(float(thr[0]) - (top.mean() + top.std(ddof=1))) < 1e-6
    assert 0.6 < float(thr[0]) < 0.8
    assert cand.tolist() == [[7, -1, -1, -1, -1]] * 2
    assert _max(scores, img) == 0.0
    # (b) top-50: low threshold -> all of the first five qualify, two of them are skipped ids: three candidates, in order
    _, cand, thr = O.seg_select(img, text, 50)
    assert float(thr[0]) < 0.5
    assert cand.tolist() == [[7, 3, 9, -1, -1]] * 2
    # (c) only the first max_cand of the top-k are ever looked at (class 11 ranks sixth here)
    img2 = img.clone()
    img2[:, 11] = 0.45
    _, cand, _ = O.seg_select(img2, text, 50)
    assert cand.tolist() == [[7, 3, 9, -1, -1]] * 2
    # normalised, nearest-up-sampled maps
    sim = torch.zeros(1, 4, 6)
    sim[0, :, 2] = torch.tensor([1.0, 2.0, 3.0, 5.0])
    maps = O.seg_norm_maps(sim, torch.tensor([[2, -1]], dtype=torch.int32), 2, 2, scale=2)
    want = np.array([[0, 0, .25, .25], [0, 0, .25, .25], [.5, .5, 1, 1], [.5, .5, 1, 1]], dtype=np.float32)
    assert maps.shape == (1, 2, 4, 4)
    assert _max(maps[0, 0], want) < 1e-7 and float(maps[0, 1].abs().max()) == 0.0


def test_seg_glue_vs_reference_fixture():
    """tests/golden/seg_glue.npz holds what the reference's OWN lines of tools/seg_evaluation.py produced
    (oracle/make_golden.py:seg_glue executes ``zero_shot_classifier`` and the per-image block of ``evaluate_benchmark``
    read from the reference tree): class embeddings, thresholds, the candidate lists (ids 0 / 255 skipped, stop below the
    threshold, only the first five of the top-k) and the min-max normalised maps."""
    z = np.load(os.path.join(GOLD, "seg_glue.npz"))
    assert _max(O.zero_shot_class_embedding(torch.tensor(z["zs_prompt"])), z["zs_weights"]) < 1e-6
    text, pooled, feats = torch.tensor(z["sel_text"]), torch.tensor(z["sel_pooled"]), torch.tensor(z["sel_feats"])
    sim, _ = O.patch_text_sim(feats, text)
    seen = set()
    for k in (10, 50):
        _, cand, thr = O.seg_select(pooled, text, k)
        maps = O.seg_norm_maps(sim, cand, 14, 14, 16)
        for b in range(pooled.shape[0]):
            assert cand[b].tolist() == z[f"sel{k}_cand_{b}"].tolist()
            assert abs(float(thr[b]) - float(z[f"sel{k}_thr_{b}"])) < 1e-6
            if f"sel{k}_map0_{b}" in z:
                assert _max(maps[b, 0, ::16, ::16], z[f"sel{k}_map0_{b}"]) < 1e-6
                assert _max(maps[b, 0], O.upsample_nearest(torch.tensor(z[f"sel{k}_map0_{b}"]), 16)) < 1e-6
            seen.update(c for c in cand[b].tolist() if c >= 0)
    assert 0 not in seen and 255 not in seen and len(seen) >= 8


def test_philox_known_answers():
    """Random123's published known-answer vectors for philox4x32-10 (kat_vectors: zeros, all ones, digits of pi)."""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = O.philox4x32_10(np.array([ctr], dtype=np.uint32), key)[0]
        assert tuple(int(x) for x in got) == want
    assert O.dropout_threshold(0.1) == 429496730 and O.dropout_threshold(0.0) == 0
    keep = O.hidden_keep_mask(7, 1, 0, 512, 768, 0.1)
    assert abs(keep.float().mean().item() - 0.9) < 3e-3
    assert not torch.equal(keep, O.hidden_keep_mask(7, 2, 0, 512, 768, 0.1))      # step, site and seed all move the stream
    assert not torch.equal(keep, O.hidden_keep_mask(7, 1, 1, 512, 768, 0.1))
    assert not torch.equal(keep, O.hidden_keep_mask(8, 1, 0, 512, 768, 0.1))


def test_bert_train_mode_dropout_vs_transformers_fixture():
    """tests/golden/bert_dropout.npz holds the output of the installed transformers BertModel in train() mode when its
    F.dropout calls are served the oracle's Philox masks in call order (make_golden.py:bert_dropout, which also checks the
    call order itself and five parameter gradients): the oracle must place its dropouts where HF does."""
    z = np.load(os.path.join(GOLD, "bert_dropout.npz"))
    seed, step = int(z["bd_seed"]), int(z["bd_step"])
    sd = O.make_state_dict(384, 6, seed=0)
    batch = O.make_batch(3, 25, seed=5)
    drop = O.PhiloxDropout(seed, step, 0.1, 0.1)
    with torch.no_grad():
        h = O.bert_forward(sd, batch["input_ids"], batch["attention_mask"], 12, O.TXT_PREFIX, dropout=drop)
        h0 = O.bert_forward(sd, batch["input_ids"], batch["attention_mask"], 12, O.TXT_PREFIX)
    ref = torch.from_numpy(z["bd_hidden"])
    assert (h[:, :, :96] - ref).abs().max().item() < 2e-4
    assert (h0[:, :, :96] - ref).abs().max().item() > 0.1                      # the masks act
    assert abs(float(z["bd_keep_rate_site1"]) - O.attn_keep_mask(seed, step, 1, 3, 12, 25, 0.1).float().mean().item()) < 1e-7
