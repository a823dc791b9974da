"""LoDA head microbenchmark (GPU box)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simseg_b200 import ops
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(4096, 197, 512, device="cuda", generator=g).bfloat16()
def t(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
print("topk_pool_l2norm_fwd 4096x196x512 k=5:", t(lambda: ops.topk_pool_l2norm_fwd(x, 5, 1, 196)), "ms")
