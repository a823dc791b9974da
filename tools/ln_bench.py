"""LayerNorm / colsum microbenchmark (GPU box): ms and achieved HBM GB/s at the ViT-S / BERT shapes of the bench step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simseg_b200 import ops


def t(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


for (M, D) in ((4096 * 197, 384), (4096 * 25, 768), (1024 * 197, 768)):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(M, D, device="cuda", generator=g)
    add = torch.randn(M, D, device="cuda", generator=g).bfloat16()
    dy = torch.randn(M, D, device="cuda", generator=g).bfloat16()
    gam, bet = torch.randn(D, device="cuda", generator=g), torch.randn(D, device="cuda", generator=g)
    s, y, _, mean, rstd = ops.add_layernorm_fwd(x, add, gam, bet, 1e-6)
    ms = t(lambda: ops.add_layernorm_fwd(x, add, gam, bet, 1e-6))
    print(f"M={M} D={D} add_ln_fwd {ms:.3f} ms {M * D * (4 + 2 + 4 + 2) / ms / 1e6:.0f} GB/s")
    dx = torch.zeros(M, D, device="cuda")
    gb = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    dg, db, dc = torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda"), torch.zeros(D, device="cuda")
    ms = t(lambda: ops.layernorm_bwd(dy, x, gam, mean, rstd, dx=dx, dx_accumulate=True, dx_bf16=gb, dgamma=dg, dbeta=db, dx_colsum=dc))
    print(f"M={M} D={D} ln_bwd     {ms:.3f} ms {M * D * (2 + 4 + 4 + 4 + 2) / ms / 1e6:.0f} GB/s")
    w = torch.randn(M, 3 * D, device="cuda", generator=g).bfloat16()
    o = torch.zeros(3 * D, device="cuda")
    ms = t(lambda: ops.colsum(w, o, accumulate=True))
    print(f"M={M} D={3 * D} colsum     {ms:.3f} ms {M * 3 * D * 2 / ms / 1e6:.0f} GB/s")
