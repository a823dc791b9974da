"""Pin the oracle against the reference and write ``tests/golden/*.npz``.

Runs ONLY in the build container (needs ``/root/reference``).  It

1. imports the reference's own modules (with ``oracle/shims`` standing in for the
   absent ``timm``) and checks every oracle function against them on seeded inputs;
2. cross-checks the ViT restatement against torchvision's independent
   ``VisionTransformer`` and the BERT restatement against the installed
   ``transformers`` ``BertModel``;
3. spawns a 2-rank gloo group to pin the global-reduce NCE branch (``GatherLayer``
   + ``targets = arange(b*rank, b*(rank+1))``) fwd and bwd;
4. stores the REFERENCE outputs (not the oracle's) as small fixtures.

    python oracle/make_golden.py            # regenerate + verify
"""
from __future__ import annotations

import os
import sys
import tempfile

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
REF = "/root/reference"
sys.path[:0] = [REF, os.path.join(HERE, "shims"), ROOT]

from oracle import simseg_oracle as O  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
TOL = 2e-5


def _check(name, a, b, tol=TOL):
    a, b = torch.as_tensor(a).double(), torch.as_tensor(b).double()
    err = (a - b).abs().max().item() if a.numel() else 0.0
    status = "ok" if err <= tol else "MISMATCH"
    print(f"  {name:48s} max|diff| = {err:.3e}  [{status}]")
    assert err <= tol, name
    return err


_CFG = {}


def _ref_cfg(yaml_name, extra=()):
    if yaml_name in _CFG:            # the reference's cfg is a global that freezes after one load
        return _CFG[yaml_name]
    import simseg.core  # noqa: F401  (must precede simseg.models)
    from simseg.core import cfg, update_cfg
    from simseg.tasks.clip.config import task_cfg_init_fn, update_clip_config
    argv = ["model.image_encoder.pretrained=False", "model.text_encoder.pretrained=False",
            "loss.global_reduce=False", "transforms.input_size=224"] + list(extra)
    cwd = os.getcwd()
    update_cfg(task_cfg_init_fn, os.path.join(REF, "configs/clip", yaml_name), argv,
               preprocess_fn=update_clip_config)
    os.chdir(cwd)
    _CFG[yaml_name] = cfg
    return cfg


def heads_and_loss():
    print("[heads / loss vs reference modules]")
    import simseg.core  # noqa: F401
    from simseg.models.components import SimpleProjection, TopKPooling, L2norm
    from simseg.tasks.clip.hooks.utils import IndexedEmbInfo, EmbANN, RetrievalMetric
    from simseg.utils.interpolate_pe import interpolate_pos_embed
    g = torch.Generator().manual_seed(7)
    out = {}
    x_img = torch.randn(6, 196, 384, generator=g)
    x_txt = torch.randn(6, 25, 768, generator=g)
    lens = torch.tensor([25, 8, 13, 1, 20, 9])
    mask = (torch.arange(25)[None] < lens[:, None]).long()
    torch.manual_seed(7)                          # nn.Linear's default init draws from the global RNG: seed it so the fixture regenerates
    pi = SimpleProjection(None, 384, 512); pt = SimpleProjection(None, 768, 512)
    with torch.no_grad():
        r_pi = pi(x_img); r_pt = pt(x_txt)
        _check("SimpleProjection img", O.simple_projection(x_img, pi.linear.weight), r_pi)
        r_pool_i = TopKPooling(5, 1)(r_pi.clone())
        r_pool_t = TopKPooling(1, 1)(r_pt.clone(), mask)
        r_pool_t3 = TopKPooling(3, 1)(r_pt.clone(), mask)          # k shrinks to min_len=1
        _check("TopKPooling k=5", O.topk_pooling(r_pi, 5), r_pool_i)
        _check("TopKPooling masked k=1", O.topk_pooling(r_pt, 1, mask), r_pool_t)
        _check("TopKPooling masked k=3->1", O.topk_pooling(r_pt, 3, mask), r_pool_t3)
        r_ni, r_nt = L2norm(r_pool_i, dim=-1), L2norm(r_pool_t, dim=-1)
        _check("L2norm", O.l2norm(r_pool_i), r_ni)
        _check("image_embed", O.image_embed(torch.cat([x_img[:, :1], x_img], 1), pi.linear.weight, 5), r_ni)
        _check("text_embed", O.text_embed(x_txt, pt.linear.weight, mask, 1), r_nt)
    out.update(heads_x_img=x_img[:2, :, :64].numpy(), heads_seed=np.array(7),
               heads_img_emb=r_ni.numpy(), heads_txt_emb=r_nt.numpy(), heads_mask=mask.numpy(),
               heads_wi=pi.linear.weight.detach().numpy(), heads_wt=pt.linear.weight.detach().numpy())

    # NCE non-global branch == W=1 global (mml_loss.py:79-87)
    from simseg.models.criteria.losses.mml_loss import NCE
    cfg = _ref_cfg("simseg.vit-s.yaml")
    nce = NCE(cfg, 0)
    a = r_ni.clone().requires_grad_(True); b = r_nt.clone().requires_grad_(True)
    loss, i2t, t2i = nce(a, b)
    loss.backward()
    a2 = r_ni.clone().requires_grad_(True); b2 = r_nt.clone().requires_grad_(True)
    temp = nce.temperature.detach().clone().requires_grad_(True)
    l2, ai, at = O.clip_loss(a2, b2, a2, b2, temp, 0)
    l2.backward()
    _check("NCE loss (W=1)", l2, loss); _check("NCE i2t acc", ai, i2t); _check("NCE t2i acc", at, t2i)
    _check("NCE d/d img", a2.grad, a.grad); _check("NCE d/d txt", b2.grad, b.grad)
    _check("NCE d/d temperature", temp.grad, nce.temperature.grad, 1e-3)
    out.update(nce_loss=loss.detach().numpy(), nce_i2t=i2t.numpy(), nce_t2i=t2i.numpy(),
               nce_dimg=a.grad.numpy(), nce_dtxt=b.grad.numpy(), nce_dtemp=nce.temperature.grad.numpy())

    # retrieval
    left = F.normalize(torch.randn(40, 512, generator=g), dim=-1)
    right = F.normalize(torch.randn(200, 512, generator=g), dim=-1)
    lg = torch.arange(40); rg = torch.arange(200) // 5
    L, R = IndexedEmbInfo("image", lg, left), IndexedEmbInfo("text", rg, right)
    _, matched = EmbANN()(L, R)
    has_r, first_r = torch.max(matched, dim=1)
    has, first = O.retrieval_first_match_rank(O.allpairs_sim(left, right), lg, rg)
    _check("retrieval first-match rank", first, first_r, 0)
    rec_r = RetrievalMetric(with_prefix=False)(L, R)
    rec = O.recall_at(has, first)
    for k in rec:
        _check(f"retrieval {k}", rec[k], rec_r[k], 1e-7)
    out.update(retr_left=left.numpy(), retr_right=right.numpy(), retr_first=first_r.numpy(),
               retr_r1=np.array(rec_r["R@1"]), retr_r5=np.array(rec_r["R@5"]), retr_r10=np.array(rec_r["R@10"]))

    # pos-embed interpolation 14x14 -> 18x18 (seg_evaluation.py:228-231)
    class _V:  # minimal visual_encoder view the reference function reads
        pass
    v = _V(); v.patch_embed = _V(); v.patch_embed.num_patches = 324; v.pos_embed = torch.zeros(1, 325, 384)
    pe = torch.randn(1, 197, 384, generator=g)
    r_pe = interpolate_pos_embed(pe, v)
    _check("interpolate_pos_embed 196->324", O.interpolate_pos_embed(pe, 324), r_pe)
    out.update(pe_in=pe[:, :, :16].numpy(), pe_out=r_pe[:, :, :16].numpy())

    # seg sim-map lines (tools/seg_evaluation.py:111-112,136 — the tool itself is a script that
    # imports pydensecrf, absent here; its two torch lines are executed verbatim on the same data)
    pf = torch.randn(2, 196, 512, generator=g)
    tf = O.zero_shot_class_embedding(torch.randn(20, 80, 512, generator=g))
    im_f_a = F.normalize(pf[0], dim=-1, p=2)
    for c in (0, 7, 19):
        attn = im_f_a @ tf[c].unsqueeze(-1)
        _check(f"seg sim map class {c}", O.patch_text_sim(pf, tf)[0][0, :, c], attn[:, 0])
    sim, am = O.patch_text_sim(pf, tf)
    out.update(seg_patch=pf.numpy(), seg_text=tf.numpy(), seg_sim=sim.numpy(), seg_argmax=am.numpy())
    return out


def encoders_cross_check():
    print("[ViT vs torchvision, BERT vs transformers]")
    out = {}
    from torchvision.models.vision_transformer import VisionTransformer
    for D, H in ((384, 6),):
        torch.manual_seed(0)
        tv = VisionTransformer(image_size=224, patch_size=16, num_layers=12, num_heads=H,
                               hidden_dim=D, mlp_dim=4 * D).eval()
        with torch.no_grad():
            for p_ in tv.parameters():
                if p_.dim() == 1:
                    p_.add_(torch.randn_like(p_) * 0.05)
            tv.class_token.normal_(std=0.02)
        sd = {"cls_token": tv.class_token, "pos_embed": tv.encoder.pos_embedding,
              "patch_embed.proj.weight": tv.conv_proj.weight, "patch_embed.proj.bias": tv.conv_proj.bias,
              "norm.weight": tv.encoder.ln.weight, "norm.bias": tv.encoder.ln.bias}
        for i in range(12):
            l = getattr(tv.encoder.layers, f"encoder_layer_{i}")
            b = f"blocks.{i}."
            sd.update({b + "norm1.weight": l.ln_1.weight, b + "norm1.bias": l.ln_1.bias,
                       b + "attn.qkv.weight": l.self_attention.in_proj_weight, b + "attn.qkv.bias": l.self_attention.in_proj_bias,
                       b + "attn.proj.weight": l.self_attention.out_proj.weight, b + "attn.proj.bias": l.self_attention.out_proj.bias,
                       b + "norm2.weight": l.ln_2.weight, b + "norm2.bias": l.ln_2.bias,
                       b + "mlp.fc1.weight": l.mlp[0].weight, b + "mlp.fc1.bias": l.mlp[0].bias,
                       b + "mlp.fc2.weight": l.mlp[3].weight, b + "mlp.fc2.bias": l.mlp[3].bias})
        sd = {k: v.detach() for k, v in sd.items()}
        img = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(3))
        with torch.no_grad():
            x = tv._process_input(img)
            x = torch.cat([tv.class_token.expand(2, -1, -1), x], dim=1)
            r = tv.encoder(x)
            o = O.vit_forward(sd, img, H)
        _check(f"ViT D={D} tokens vs torchvision", o, r, 2e-4)

    from transformers import BertConfig, BertModel
    torch.manual_seed(0)
    bm = BertModel(BertConfig(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0),
                   add_pooling_layer=False).eval()
    bsd = {k: v.detach() for k, v in bm.state_dict().items()}
    batch = O.make_batch(3, 25, seed=5)
    with torch.no_grad():
        r = bm(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"]).last_hidden_state
        o = O.bert_forward(bsd, batch["input_ids"], batch["attention_mask"])
    _check("BERT last_hidden_state vs transformers", o, r, 2e-4)
    return out


def bert_dropout():
    """Pins WHERE the oracle applies BERT's train-mode dropout (sites and order, oracle/simseg_oracle.py) against the installed
    ``transformers`` BertModel in train() mode: ``torch.nn.functional.dropout`` is replaced for the duration of one forward +
    backward by a function that serves the oracle's own Philox masks in call order, so both sides drop the same elements and
    the outputs / gradients must agree.  (The reference reaches this code as ``huggingface_builder.py:16-17`` under
    ``model.train()``; bert-base-uncased: hidden_dropout_prob = attention_probs_dropout_prob = 0.1.)"""
    print("[BERT train-mode dropout placement vs transformers BertModel.train()]")
    from transformers import BertConfig, BertModel
    seed, step, B, T = 20260117, 3, 3, 25
    sd_all = O.make_state_dict(384, 6, seed=0)
    bsd = {k[len(O.TXT_PREFIX):]: v for k, v in sd_all.items() if k.startswith(O.TXT_PREFIX)}
    try:
        bm = BertModel(BertConfig(hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1, attn_implementation="eager"),
                       add_pooling_layer=False)
    except TypeError:
        bm = BertModel(BertConfig(hidden_dropout_prob=0.1, attention_probs_dropout_prob=0.1), add_pooling_layer=False)
    missing, unexpected = bm.load_state_dict(bsd, strict=False)
    assert not unexpected and all("position_ids" in m or "token_type_ids" in m for m in missing), (missing, unexpected)
    bm.train()
    batch = O.make_batch(B, T, seed=5)
    drop = O.PhiloxDropout(seed, step, 0.1, 0.1)
    calls = []
    real = F.dropout

    def served(x, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return x
        site = len(calls)
        calls.append(tuple(x.shape))
        assert abs(p - 0.1) < 1e-12
        return drop.attn(site, x) if x.dim() == 4 else drop.hidden(site, x)

    torch.nn.functional.dropout = served
    try:
        r = bm(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"]).last_hidden_state
        g = torch.randn(r.shape, generator=torch.Generator().manual_seed(9))
        (r * g).sum().backward()
    finally:
        torch.nn.functional.dropout = real
    want = [(B, T, 768)]
    for _ in range(12):
        want += [(B, 12, T, T), (B, T, 768), (B, T, 768)]
    assert calls == want, "HF's dropout call order differs from the oracle's site numbering"
    print("  dropout call order: embeddings, then 12 x (attention probs, attention.output, output)  [ok]")
    sdg = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in bsd.items()}
    o = O.bert_forward(sdg, batch["input_ids"], batch["attention_mask"], dropout=drop)
    (o * g).sum().backward()
    _check("BERT train() last_hidden_state, served masks", o, r, 2e-4)
    named = dict(bm.named_parameters())
    for k in ("encoder.layer.0.attention.self.query.weight", "encoder.layer.5.attention.output.dense.bias",
              "encoder.layer.11.output.dense.weight", "embeddings.word_embeddings.weight", "embeddings.LayerNorm.weight"):
        rg, og = named[k].grad, sdg[k].grad
        sc = rg.abs().max().item() + 1e-12
        _check(f"grad {k[-44:]}", og / sc, rg / sc, 2e-3)
    # eval-mode output differs (the masks really act), and the p = 0 oracle reproduces it
    bm.eval()
    with torch.no_grad():
        r0 = bm(input_ids=batch["input_ids"], attention_mask=batch["attention_mask"]).last_hidden_state
    assert (r0 - r).abs().max().item() > 0.1
    return dict(bd_seed=np.array(seed), bd_step=np.array(step), bd_hidden=r.detach()[:, :, :96].contiguous().numpy(),
                bd_grad_g=g[:, :, :8].contiguous().numpy(),
                bd_dq0_norm=np.array(named["encoder.layer.0.attention.self.query.weight"].grad.norm().item()),
                bd_keep_rate_site1=np.array(O.attn_keep_mask(seed, step, 1, B, 12, T, 0.1).float().mean().item()))


def full_model():
    print("[reference CLIPModel (ViT-S + BERT-base) vs oracle, same weights]")
    import simseg.core  # noqa: F401
    from simseg.models import PIPELINE
    from simseg.utils import build_from_cfg
    tmp = tempfile.mkdtemp()
    from transformers import BertConfig
    BertConfig(hidden_dropout_prob=0.0, attention_probs_dropout_prob=0.0).save_pretrained(
        os.path.join(tmp, "bert-base-uncased"))
    cwd = os.getcwd(); os.chdir(tmp)
    try:
        cfg = _ref_cfg("simseg.vit-s.yaml")
        model = build_from_cfg(cfg.model.name, cfg, PIPELINE)
    finally:
        os.chdir(cwd)
    sd = O.make_state_dict(384, 6, seed=0)
    missing, unexpected = model.load_state_dict(sd, strict=False)
    missing = [m for m in missing if "pooler" not in m and "position_ids" not in m and "token_type_ids" not in m]
    assert not missing and not unexpected, (missing, unexpected)
    print("  state-dict keys: reference accepts all", len(sd), "oracle keys (SURVEY §8b naming)")
    model.eval()                                  # dropout off; parity is defined at p=0
    B = 8
    batch = O.make_batch(B, 25, seed=1234)
    ref_batch = {k: v.clone() for k, v in batch.items()}
    model.zero_grad()
    loss_dict, i2t, t2i = model(ref_batch)
    loss = loss_dict["nce_loss"]
    loss.backward()
    with torch.no_grad():
        img_e, txt_e = model({k: v.clone() for k, v in batch.items()}, embeddings="all")
        tok = model.forward_image_feature(batch["image"])
    sdg = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    l_o, ai, at = O.clip_train_forward(sdg, batch, 6)
    l_o.backward()
    oi, ot = O.clip_embeddings(sd, batch, 6)
    _check("CLIP img embeddings", oi, img_e, 1e-4); _check("CLIP txt embeddings", ot, txt_e, 1e-4)
    _check("CLIP loss", l_o, loss, 1e-4); _check("CLIP i2t acc", ai, i2t); _check("CLIP t2i acc", at, t2i)
    named = dict(model.named_parameters())
    gkeys = ["image_projection.linear.weight", "text_projection.linear.weight", "loss.temperature",
             O.IMG_PREFIX + "blocks.0.attn.qkv.weight", O.IMG_PREFIX + "patch_embed.proj.weight",
             O.IMG_PREFIX + "pos_embed", O.IMG_PREFIX + "blocks.11.mlp.fc2.bias",
             O.TXT_PREFIX + "encoder.layer.0.attention.self.query.weight",
             O.TXT_PREFIX + "encoder.layer.11.output.LayerNorm.weight",
             O.TXT_PREFIX + "embeddings.position_embeddings.weight"]
    out = {}
    for k in gkeys:
        rg, og = named[k].grad, sdg[k].grad
        scale = rg.abs().max().item() + 1e-12
        _check(f"grad {k[-44:]}", og / scale, rg / scale, 2e-3)
        out["grad_norm/" + k] = np.array(rg.norm().item())
    out.update(clip_img_emb=img_e.numpy(), clip_txt_emb=txt_e.numpy(), clip_loss=loss.detach().numpy(),
               clip_i2t=i2t.numpy(), clip_t2i=t2i.numpy(),
               clip_tokens_head=tok[:, :4, :32].contiguous().numpy(),
               clip_dWi_head=named["image_projection.linear.weight"].grad[:8, :32].numpy(),
               clip_dtemp=named["loss.temperature"].grad.numpy())
    return out


def _gloo_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), HOSTNAME="box",
                      RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    sys.path[:0] = [REF, os.path.join(HERE, "shims"), ROOT]
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from simseg.utils import ENV
    from simseg.utils.dist import GatherLayer
    g = torch.Generator().manual_seed(11)
    img = O.l2norm(torch.randn(world * 6, 512, generator=g))
    txt = O.l2norm(torch.randn(world * 6, 512, generator=g))
    b = 6
    li = img[rank * b:(rank + 1) * b].clone().requires_grad_(True)
    lt = txt[rank * b:(rank + 1) * b].clone().requires_grad_(True)
    temp = torch.tensor(0.02)
    # the reference's global branch, mml_loss.py:58-77,89-95, with its own GatherLayer
    def ref_dir(f1, f2):
        f2g = GatherLayer.apply(f2, dist.group.WORLD, rank)
        logits = (f1 @ f2g.T) / torch.clamp(temp, 0.001, 0.5)
        targets = torch.arange(b * rank, b * (rank + 1))
        return F.cross_entropy(logits, targets, reduction="none").mean()
    loss = 0.5 * (ref_dir(li, lt) + ref_dir(lt, li))
    loss.backward()
    q.put((rank, loss.item(), li.grad.numpy(), lt.grad.numpy(), img.numpy(), txt.numpy()))
    dist.barrier(); dist.destroy_process_group()


def global_reduce_two_ranks():
    print("[global-reduce NCE with the reference GatherLayer, 2 gloo ranks]")
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29541
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = sorted([q.get(timeout=300) for _ in ps], key=lambda t: t[0])
    [p.join() for p in ps]
    out = {}
    img, txt = torch.tensor(res[0][4]), torch.tensor(res[0][5])
    for rank, loss, gi, gt, _, _ in res:
        b = 6
        ig = img.clone().requires_grad_(True); tg = txt.clone().requires_grad_(True)
        # per-rank loss as the oracle states it; gradient wrt a rank's local rows is the SUM over
        # ranks of d loss_r (GatherLayer.backward all_reduce, utils/dist.py:348-354)
        total = 0
        for r in range(2):
            l, _, _ = O.clip_loss(ig[r * b:(r + 1) * b], tg[r * b:(r + 1) * b], ig, tg, torch.tensor(0.02), r)
            if r == rank:
                _check(f"rank {rank} loss", l, loss, 1e-5)
            total = total + l
        total.backward()
        _check(f"rank {rank} d img (sum over ranks)", ig.grad[rank * b:(rank + 1) * b], gi, 1e-5)
        _check(f"rank {rank} d txt (sum over ranks)", tg.grad[rank * b:(rank + 1) * b], gt, 1e-5)
        out[f"gr_loss_{rank}"] = np.array(loss); out[f"gr_dimg_{rank}"] = gi; out[f"gr_dtxt_{rank}"] = gt
    out["gr_img"] = img.numpy(); out["gr_txt"] = txt.numpy()
    return out


def seg_glue():
    """Pin the zero-shot segmentation glue against the reference's OWN source lines.  ``tools/seg_evaluation.py`` is a
    script that does not import here (pydensecrf, cv2 ...), so its code is read from /root/reference at run time and executed
    piecewise: the whole ``zero_shot_classifier`` function (AST-extracted) with stub model / tokenizer objects, and the
    per-image block of ``evaluate_benchmark`` from ``im_f_a = F.normalize(...)`` down to ``norm_attn = ...`` (the statements
    before the CRF call), with a recording line appended to the candidate loop.  Nothing of it is copied into the repo."""
    print("[zero-shot segmentation glue vs the lines of tools/seg_evaluation.py]")
    import ast
    import textwrap
    src = open(os.path.join(REF, "tools", "seg_evaluation.py")).read()
    out = {}
    g = torch.Generator().manual_seed(21)

    # ---- zero_shot_classifier(model, classnames, make_template, tokenizer, ENV)
    fn = next(n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "zero_shot_classifier")
    ns = {"torch": torch, "tqdm": lambda x: x}
    exec(compile(ast.Module(body=[fn], type_ignores=[]), "seg_evaluation.py", "exec"), ns)
    Cn, P, T, E = 7, 5, 25, 64
    prompt_emb = torch.randn(Cn, P, E, generator=g)

    class _Text:                                   # stands in for model.module: returns the prepared prompt embeddings
        def forward_text_feature(self, input_ids, attention_mask):
            assert input_ids.shape == (P, T) and attention_mask.shape == (P, T)
            return int(input_ids[0, 0])
        def forward_text_project(self, feat, attention_mask):
            return prompt_emb[feat].clone()
    model = type("M", (), {"module": _Text()})()
    tok = lambda texts, **kw: {"input_ids": [[int(texts[0])] * T for _ in texts], "attention_mask": [[1] * T for _ in texts]}
    env = type("E", (), {"rank": "cpu"})()
    cuda_attr = torch.Tensor.cuda
    torch.Tensor.cuda = lambda self, *a, **k: self                    # the function ends with .cuda(); there is no GPU here
    try:
        ref_w = ns["zero_shot_classifier"](model, list(range(Cn)), lambda c: [str(c)] * P, tok, env)
    finally:
        torch.Tensor.cuda = cuda_attr
    _check("zero_shot_classifier -> class embeddings", O.zero_shot_class_embedding(prompt_emb), ref_w, 1e-6)
    out.update(zs_prompt=prompt_emb.numpy(), zs_weights=ref_w.numpy())

    # ---- per-image block of evaluate_benchmark
    lines = src.splitlines()
    i0 = next(i for i, l in enumerate(lines) if "im_f_a = F.normalize(im_f_a" in l)
    i1 = next(i for i, l in enumerate(lines) if "norm_attn = (attn_ai2at - min_value)" in l)
    block = lines[i0:i1 + 1]
    indent = len(lines[i1]) - len(lines[i1].lstrip())
    block.append(" " * indent + "_rec.append((int(index), norm_attn.copy()))")
    code = compile(textwrap.dedent("\n".join(block)), "seg_evaluation.py[block]", "exec")
    B, hw, Cc, Ee = 4, 14, 300, 48
    text = F.normalize(torch.randn(Cc, Ee, generator=g), dim=-1)
    feats = torch.randn(B, hw * hw, Ee, generator=g)
    favored = [[0, 7, 255, 3, 9, 11], [5, 6, 8, 10, 12, 13], [255, 0, 20, 21, 22, 23], [40]]
    pooled = []
    for b in range(B):
        v = 0.05 * torch.randn(Ee, generator=g)
        for rank_, c in enumerate(favored[b]):
            v = v + (1.0 - 0.12 * rank_) * text[c]
        pooled.append(F.normalize(v, dim=-1))
    pooled = torch.stack(pooled)
    for top_cls_num in (10, 50):
        o_scores, o_cand, o_thr = O.seg_select(pooled, text, top_cls_num)
        sim, _ = O.patch_text_sim(feats, text)
        o_maps = O.seg_norm_maps(sim, o_cand, hw, hw, 16)
        for b in range(B):
            rec = []
            nsb = {"torch": torch, "np": np, "F": F, "im_f_a": feats[b].clone(), "image_feature_pooled": pooled, "index": b,
                   "label_text_feature": text, "top_cls_num": top_cls_num, "num_patch": hw, "patch_size": 16,
                   "image_shape": (hw * 16, hw * 16), "seg_categories": list(range(Cc)), "_rec": rec}
            exec(code, nsb)
            _check(f"top{top_cls_num} img {b}: scores", o_scores[b], nsb["scores"], 1e-6)
            _check(f"top{top_cls_num} img {b}: threshold", o_thr[b], nsb["threshold"], 1e-6)
            ref_c = [c for c, _ in rec] + [-1] * (5 - len(rec))
            assert o_cand[b].tolist() == ref_c, (o_cand[b].tolist(), ref_c)
            for k, (_, m) in enumerate(rec):
                _check(f"top{top_cls_num} img {b}: map {k}", o_maps[b, k], m, 1e-6)
            out[f"sel{top_cls_num}_cand_{b}"] = np.array(ref_c, dtype=np.int32)
            out[f"sel{top_cls_num}_thr_{b}"] = np.array(float(nsb["threshold"]), dtype=np.float32)
            if rec:
                out[f"sel{top_cls_num}_map0_{b}"] = rec[0][1][::16, ::16].astype(np.float32)
        print(f"  top{top_cls_num}: candidates per image", [[c for c in o_cand[b].tolist() if c >= 0] for b in range(B)])
    out.update(sel_text=text.numpy(), sel_pooled=pooled.numpy(), sel_feats=feats.numpy())
    return out


def main():
    torch.set_num_threads(os.cpu_count())
    os.makedirs(GOLD, exist_ok=True)
    np.savez_compressed(os.path.join(GOLD, "heads_loss.npz"), **heads_and_loss())
    encoders_cross_check()
    np.savez_compressed(os.path.join(GOLD, "global_reduce.npz"), **global_reduce_two_ranks())
    np.savez_compressed(os.path.join(GOLD, "clip_vit_s.npz"), **full_model())
    np.savez_compressed(os.path.join(GOLD, "seg_glue.npz"), **seg_glue())
    np.savez_compressed(os.path.join(GOLD, "bert_dropout.npz"), **bert_dropout())
    print("golden vectors written to", GOLD)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "seg_glue":          # regenerate this fixture alone
        np.savez_compressed(os.path.join(GOLD, "seg_glue.npz"), **seg_glue())
    elif len(sys.argv) > 1 and sys.argv[1] == "bert_dropout":
        np.savez_compressed(os.path.join(GOLD, "bert_dropout.npz"), **bert_dropout())
    else:
        main()
