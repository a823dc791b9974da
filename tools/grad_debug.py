"""GPU debug aid (test infrastructure, not product): per-parameter gradient error of the drop-in CLIPModel against the
fp32 oracle for fixed embedding cotangents, next to the error the oracle itself shows when run under
torch.autocast(bfloat16) (the reference's precision contract) — separates kernel bugs from bf16 conditioning."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import simseg_oracle as O  # noqa: E402
from simseg_b200.config import load_cfg  # noqa: E402
from simseg_b200.pipeline import PIPELINE  # noqa: E402


def main():
    dev = torch.device("cuda:0")
    torch.backends.cuda.matmul.allow_tf32 = False
    cfg = load_cfg("simseg.vit-s.yaml", ["model.image_encoder.pretrained=False", "model.text_encoder.pretrained=False",
                                          "transforms.input_size=224"])
    model = PIPELINE["clip"](cfg).to(dev)
    sd = O.make_state_dict(384, 6, seed=0)
    model.load_state_dict(sd, strict=True)
    B = int(os.environ.get("B", "8"))
    batch = O.make_batch(B, 25, seed=1234)
    gb = {k: v.to(dev) for k, v in batch.items()}
    g = torch.Generator().manual_seed(99)
    gi, gt = torch.randn(B, 512, generator=g).to(dev), torch.randn(B, 512, generator=g).to(dev)
    model.zero_grad(set_to_none=True)
    ie, te = model(gb, embeddings="all")
    torch.autograd.backward([ie, te], [gi, gt])
    ours = {k: p.grad.detach().clone() for k, p in model.named_parameters() if p.grad is not None}

    def oracle_grads(autocast):
        sdg = {k: v.to(dev).clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
        with torch.autocast("cuda", dtype=torch.bfloat16, enabled=autocast):
            oi, ot = O.clip_embeddings(sdg, gb, 6)
        torch.autograd.backward([oi.float(), ot.float()], [gi, gt])
        return {k: v.grad for k, v in sdg.items() if v.grad is not None}, oi.detach().float(), ot.detach().float()

    ref, ri, rt = oracle_grads(False)
    amp, ai, at = oracle_grads(True)
    print("emb max|diff| ours-fp32: img %.2e txt %.2e | autocast-fp32: img %.2e txt %.2e" % (
        (ie - ri).abs().max().item(), (te - rt).abs().max().item(), (ai - ri).abs().max().item(), (at - rt).abs().max().item()))
    rel = lambda a, b: ((a - b).norm() / (b.norm() + 1e-30)).item()
    print(f"{'parameter':84s} ours/fp32  amp/fp32")
    for k in ours:
        if k in ref and ref[k].norm() > 1e-7:
            print(f"{k:84s} {rel(ours[k], ref[k]):8.4f}  {rel(amp[k], ref[k]):8.4f}")


if __name__ == "__main__":
    main()
