"""One launch of attention fwd / bwd per shape for `ncu --set full -k regex:attention`."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simseg_b200 import ops
g = torch.Generator(device="cuda").manual_seed(1)
for (B, H, S) in ((2048, 6, 197),):
    D = H * 64
    qkv = torch.randn(B, S, 3, H, 64, device="cuda", generator=g).bfloat16()
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    strides = (S * 3 * D, 3 * D, 64)
    dout = torch.randn(B, S, D, device="cuda", generator=g).bfloat16()
    dqkv = torch.empty_like(qkv)
    out, lse = ops.attention_fwd(q, k, v, B, H, S, strides, None, 0.125)
    ops.attention_bwd(q, k, v, out, dout, lse, B, H, S, strides, None, 0.125, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2])
    torch.cuda.synchronize()
