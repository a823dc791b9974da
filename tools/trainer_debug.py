"""Diagnose Trainer.step vs oracle divergence at step 2: stale bf16 weight copies or genuine sensitivity?"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from oracle import simseg_oracle as O
from simseg_b200.config import load_cfg
from simseg_b200.pipeline import PIPELINE
from simseg_b200.train import Trainer

cuda = torch.device("cuda:0")
cfg = load_cfg("simseg.vit-s.yaml", ["model.image_encoder.pretrained=False", "model.text_encoder.pretrained=False", "transforms.input_size=224"])
sd = O.make_state_dict(384, 6, seed=0)
batch = O.make_batch(8, 25, seed=1234)
gb = {k: v.to(cuda) for k, v in batch.items()}
model = PIPELINE["clip"](cfg).to(cuda)
model.load_state_dict(sd)
tr = Trainer(model, cfg)
l0 = tr.step(gb)[0].item()
# weights after one step
sd1 = {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}
with torch.no_grad():
    l1_same_model = model(gb)[0]["nce_loss"].item()
fresh = PIPELINE["clip"](cfg).to(cuda)
fresh.load_state_dict(sd1)
with torch.no_grad():
    l1_fresh = fresh(gb)[0]["nce_loss"].item()
l1_oracle_on_ours = O.clip_train_forward(sd1, batch, 6)[0].item()
print("step0 loss", l0, "| after 1 step: same model", l1_same_model, "fresh model", l1_fresh, "oracle on OUR weights", l1_oracle_on_ours)
# oracle's own step
op = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
p = cfg.optim.param
oopt = torch.optim.AdamW([v for v in op.values() if v.requires_grad], lr=cfg.optim.lr.init, betas=tuple(p["betas"]), eps=p["eps"], weight_decay=p["weight_decay"])
lo = O.clip_train_forward(op, batch, 6)[0]
lo.backward()
oopt.step()
sdo = {k: v.detach().clone() for k, v in op.items()}
l1_oracle = O.clip_train_forward(sdo, batch, 6)[0].item()
fresh.load_state_dict(sdo)
with torch.no_grad():
    l1_ours_on_oracle = fresh(gb)[0]["nce_loss"].item()
print("oracle step0", lo.item(), "oracle after its own step", l1_oracle, "| OUR forward on ORACLE weights", l1_ours_on_oracle)
# how different are the two updated weight sets, relative to the update size
num = sum(((sd1[k] - sdo[k]) ** 2).sum().item() for k in sdo if sdo[k].is_floating_point())
den = sum(((sdo[k] - sd[k]) ** 2).sum().item() for k in sdo if sdo[k].is_floating_point())
print("||w_ours - w_oracle|| / ||update|| =", (num / den) ** 0.5)
for k in ["loss.temperature"]:
    print(k, sd[k].item(), sd1[k].item(), sdo[k].item())
# interpolate: loss along the oracle update direction (fp32 oracle) at 0.25 steps
for a in (0.25, 0.5, 0.75, 1.0, 1.25):
    sda = {k: (sd[k] + a * (sdo[k] - sd[k])) if sd[k].is_floating_point() else sd[k] for k in sd}
    print("alpha", a, "oracle loss", O.clip_train_forward(sda, batch, 6)[0].item())
