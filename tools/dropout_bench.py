"""Cost of BERT's train-mode dropout on the B200 path (GPU box): attention fwd / bwd with and without the probability
mask, the mask generator, and the two fused LayerNorm forms, at the per-layer shapes of the bench step (b = 4096, T = 25)
and at T = 77.  Per-step cost = 12 layers x (mask + d fwd + mask + d bwd) + 25 LayerNorm pairs."""
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


def attn(B, H, S):
    g = torch.Generator(device="cuda").manual_seed(1)
    D = H * 64
    qkv = torch.randn(B, S, 3, H, 64, device="cuda", generator=g).bfloat16()
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    strides = (S * 3 * D, 3 * D, 64)
    klen = torch.randint(8, S + 1, (B,), device="cuda", generator=g, dtype=torch.int32)
    dout = torch.randn(B, S, D, device="cuda", generator=g).bfloat16()
    dqkv = torch.empty_like(qkv)
    rng = torch.tensor([123, 1], dtype=torch.int64, device="cuda")
    d = ops.Drop(0.1, rng, 1)
    mask = ops.attn_dropout_mask(B, H, S, d)
    out, lse = ops.attention_fwd(q, k, v, B, H, S, strides, klen, 0.125)
    f0 = t(lambda: ops.attention_fwd(q, k, v, B, H, S, strides, klen, 0.125, out=out, lse=lse))
    f1 = t(lambda: ops.attention_fwd(q, k, v, B, H, S, strides, klen, 0.125, out=out, lse=lse, drop_mask=mask, drop_p=0.1))
    b0 = t(lambda: ops.attention_bwd(q, k, v, out, dout, lse, B, H, S, strides, klen, 0.125, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2]))
    b1 = t(lambda: ops.attention_bwd(q, k, v, out, dout, lse, B, H, S, strides, klen, 0.125, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2],
                                     drop_mask=mask, drop_p=0.1))
    m = t(lambda: ops.attn_dropout_mask(B, H, S, d, out=mask))
    print(f"attention B={B} H={H} S={S}: fwd {f0:.3f} -> {f1:.3f} ms, bwd {b0:.3f} -> {b1:.3f} ms, mask draw {m:.3f} ms "
          f"({mask.numel() * 4 / 1e6:.1f} MB)", flush=True)


def ln(M, D):
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(M, D, device="cuda", generator=g)
    add = torch.randn(M, D, device="cuda", generator=g).bfloat16()
    gam, bet = torch.ones(D, device="cuda"), torch.zeros(D, device="cuda")
    rng = torch.tensor([123, 1], dtype=torch.int64, device="cuda")
    d = ops.Drop(0.1, rng, 2)
    s, yb, yf, mean, rstd = ops.add_layernorm_fwd(x, add, gam, bet, 1e-12, want_f32=True)
    f0 = t(lambda: ops.add_layernorm_fwd(x, add, gam, bet, 1e-12, want_f32=True))
    f1 = t(lambda: ops.add_layernorm_fwd(x, add, gam, bet, 1e-12, want_f32=True, drop=d))
    dy = torch.randn(M, D, device="cuda", generator=g).bfloat16()
    dy2 = torch.randn(M, D, device="cuda", generator=g)
    dx = torch.empty(M, D, device="cuda")
    dxb = torch.empty(M, D, device="cuda", dtype=torch.bfloat16)
    dg, db, dc = (torch.zeros(D, device="cuda") for _ in range(3))
    kw = dict(dy2=dy2, dx=dx, dx_bf16=dxb, dgamma=dg, dbeta=db, dx_colsum=dc)
    b0 = t(lambda: ops.layernorm_bwd(dy, s, gam, mean, rstd, **kw))
    b1 = t(lambda: ops.layernorm_bwd(dy, s, gam, mean, rstd, drop=d, drop_mode=1, **kw))
    fb = M * D * (4 + 2 + 4 + 2 + 4) / 1e9
    bb = M * D * (2 + 4 + 4 + 4 + 2) / 1e9
    print(f"add+LayerNorm M={M} D={D}: fwd {f0:.3f} -> {f1:.3f} ms ({fb / f0 * 1e3:.0f} -> {fb / f1 * 1e3:.0f} GB/s), "
          f"bwd {b0:.3f} -> {b1:.3f} ms ({bb / b0 * 1e3:.0f} -> {bb / b1 * 1e3:.0f} GB/s)", flush=True)


if __name__ == "__main__":
    attn(4096, 12, 25)
    attn(4096, 12, 77)
    attn(512, 12, 25)
    ln(4096 * 25, 768)
    ln(512 * 25, 768)
