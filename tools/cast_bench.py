"""Per-step bf16 refresh of every Linear weight (one simseg_cast_bf16_multi launch): time and bytes.  GPU box."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simseg_b200.config import load_cfg
from simseg_b200.pipeline import PIPELINE
from simseg_b200.synthetic import make_batch

cfg = load_cfg("simseg.vit-s.yaml", ["model.image_encoder.pretrained=False", "model.text_encoder.pretrained=False", "transforms.input_size=224"])
model = PIPELINE["clip"](cfg).cuda()
b = {k: v.cuda() for k, v in make_batch(4, 25, seed=1).items()}
model(b)[0]["nce_loss"].backward()          # registers every forward / transposed copy
wc = model._shared.wc
first = next(iter(wc._items.values()))["p"]
nbytes = sum(it["rows"] * it["cols"] * (4 + 2 * ((it["dst"] is not None) + (it["dst_t"] is not None))) for it in wc._items.values())
ts = []
for i in range(12):
    wc.clear()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    wc._refresh()
    e1.record()
    torch.cuda.synchronize()
    ts.append(e0.elapsed_time(e1))
ms = sorted(ts)[len(ts) // 2]
print(f"bf16 weight refresh: {len(wc._items)} items, {nbytes / 1e6:.0f} MB moved, {ms * 1e3:.0f} us = {nbytes / ms / 1e6:.0f} GB/s")
