"""``graph.GraphedCall`` / ``graph.graphed_segment``: a forward recorded once and replayed as one CUDA-graph launch must give
exactly what the eager call gives, for inputs other than the ones it was recorded with (the eval tools' per-batch body,
``tools/seg_evaluation.py:84-150``, and the BASELINE configs[0] forward)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _model(cuda):
    from oracle import simseg_oracle as O
    from simseg_b200.config import load_cfg
    from simseg_b200.pipeline import PIPELINE
    cfg = load_cfg("simseg.vit-s.yaml", ["model.image_encoder.pretrained=False", "model.text_encoder.pretrained=False",
                                          "transforms.input_size=224"])
    model = PIPELINE["clip"](cfg).to(cuda).eval()
    model.load_state_dict(O.make_state_dict(384, 6, seed=0))
    return model


def test_graphed_forward_equals_eager_bitwise(cuda):
    from oracle import simseg_oracle as O
    from simseg_b200 import ops
    from simseg_b200.graph import GraphedCall
    model = _model(cuda)
    text = torch.nn.functional.normalize(torch.randn(20, 512, generator=torch.Generator().manual_seed(5)), dim=-1).to(cuda)

    def fwd(image, input_ids, attention_mask):
        feat = model.forward_image_feature(image)
        img = model.forward_image_project(feat)
        sim, am = ops.patch_text_sim(model.image_projection(feat).contiguous(), text)
        txt = model.forward_text_project(model.forward_text_feature(input_ids, attention_mask), attention_mask)
        return {"sim": sim, "argmax": am, "img": img, "txt": txt}

    b0 = {k: v.to(cuda) for k, v in O.make_batch(4, 77, seed=1).items()}
    gc = GraphedCall(fwd, b0["image"], b0["input_ids"], b0["attention_mask"])
    assert gc.launches_per_replay > 50
    for seed in (2, 3):
        b = {k: v.to(cuda) for k, v in O.make_batch(4, 77, seed=seed).items()}
        with torch.no_grad():
            want = {k: v.clone() for k, v in fwd(b["image"], b["input_ids"], b["attention_mask"]).items()}
        got = gc(b["image"], b["input_ids"], b["attention_mask"])
        torch.cuda.synchronize()
        for k in want:
            assert torch.equal(got[k], want[k]), k
    # pinned host inputs are copied host-to-device by the call itself
    hb = {k: v.pin_memory() for k, v in O.make_batch(4, 77, seed=2).items()}
    got = gc(hb["image"], hb["input_ids"], hb["attention_mask"])
    torch.cuda.synchronize()
    b = {k: v.to(cuda) for k, v in hb.items()}
    with torch.no_grad():
        want = fwd(b["image"], b["input_ids"], b["attention_mask"])
    assert torch.equal(got["sim"], want["sim"]) and torch.equal(got["txt"], want["txt"])
    with pytest.raises(ValueError):
        gc(b["image"][:2], b["input_ids"], b["attention_mask"])


def test_graphed_segment_equals_segment(cuda):
    from simseg_b200 import seg
    from simseg_b200.graph import graphed_segment
    model = _model(cuda)
    g = torch.Generator(device=cuda).manual_seed(3)
    class_emb = torch.nn.functional.normalize(torch.randn(21, 512, device=cuda, generator=g), dim=-1)
    imgs = [torch.randn(2, 3, 224, 224, device=cuda, generator=g) for _ in range(3)]
    gs = graphed_segment(model, class_emb, imgs[0], top_cls_num=3)
    for x in imgs[1:]:
        want = [t.clone() for t in seg.segment(model, x, class_emb, 3)]
        got = gs(x)
        torch.cuda.synchronize()
        for a, b in zip(got, want):
            assert torch.equal(a, b)
