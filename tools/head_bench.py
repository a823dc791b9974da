"""LoDA head microbenchmark (GPU box): fused projection + top-k (simseg_proj_topk_*) vs the two-kernel path, fwd and bwd."""
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


for (tag, B, S, D, E, k, t0) in (("ViT-S image head", 4096, 197, 384, 512, 5, 1), ("ViT-B image head", 1024, 197, 768, 512, 5, 1),
                                   ("BERT text head T=25", 4096, 25, 768, 512, 1, 1)):
    g = torch.Generator(device="cuda").manual_seed(0)
    x = (torch.randn(B, S, D, device="cuda", generator=g) * 0.7).bfloat16()
    w = (torch.randn(E, D, device="cuda", generator=g) * 0.05).bfloat16()
    wt = w.t().contiguous()
    demb = torch.randn(B, E, device="cuda", generator=g)
    dw = torch.zeros(E, D, device="cuda")
    nt = S - t0

    def two_fwd():
        p = ops.linear_fwd(x.view(B * S, D), w).view(B, S, E)
        return ops.topk_pool_l2norm_fwd(p, k, t0, nt)
    pooled, emb, idx = ops.proj_topk_fwd(x, w, k, t0, nt)

    def two_bwd():
        dy = ops.topk_pool_l2norm_bwd(demb, pooled, idx, S, k)
        ops.linear_dgrad(dy.view(B * S, E), wt, out_dtype=torch.float32)
        ops.linear_wgrad(dy.view(B * S, E), x.view(B * S, D), dw, accumulate=True)
    f2, f1 = t(two_fwd), t(lambda: ops.proj_topk_fwd(x, w, k, t0, nt))
    b2, b1 = t(two_bwd), t(lambda: ops.proj_topk_bwd(demb, pooled, idx, x, wt, k, dw=dw))
    bd = t(lambda: ops.proj_topk_bwd(demb, pooled, idx, x, wt, k, dw=None))
    bw = t(lambda: ops.proj_topk_bwd(demb, pooled, idx, x, None, k, dw=dw, want_dx=False))
    print(f"{tag:22s} B={B} S={S} D={D} k={k}: fwd two-kernel {f2:.3f} ms | fused {f1:.3f} ms || bwd two-kernel {b2:.3f} ms | "
          f"fused {b1:.3f} ms (dgrad only {bd:.3f}, wgrad only {bw:.3f})", flush=True)
