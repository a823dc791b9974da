"""One launch per configuration of the fused patch-text kernel for `ncu --set full -k regex:patch_sim`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simseg_b200 import ops
g = torch.Generator(device="cuda").manual_seed(3)
for (B, C) in ((4096, 171), (4096, 20), (64, 171)):
    p = torch.randn(B * 196, 512, device="cuda", generator=g).bfloat16()
    t = torch.nn.functional.normalize(torch.randn(C, 512, device="cuda", generator=g), dim=-1).bfloat16()
    junk = torch.empty(256 << 20, device="cuda", dtype=torch.uint8).fill_(1)       # flush L2
    ops.patch_text_sim(p, t)
    torch.cuda.synchronize()
