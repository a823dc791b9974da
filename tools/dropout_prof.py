import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simseg_b200 import ops
B, H, S = 4096, 12, int(sys.argv[1]) if len(sys.argv) > 1 else 25
g = torch.Generator(device="cuda").manual_seed(1)
D = H * 64
qkv = torch.randn(B, S, 3, H, 64, device="cuda", generator=g).bfloat16()
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
strides = (S * 3 * D, 3 * D, 64)
klen = torch.randint(8, S + 1, (B,), device="cuda", generator=g, dtype=torch.int32)
dout = torch.randn(B, S, D, device="cuda", generator=g).bfloat16()
dqkv = torch.empty_like(qkv)
rng = torch.tensor([123, 1], dtype=torch.int64, device="cuda")
mask = ops.attn_dropout_mask(B, H, S, ops.Drop(0.1, rng, 1))
out, lse = ops.attention_fwd(q, k, v, B, H, S, strides, klen, 0.125, drop_mask=mask, drop_p=0.1)
ops.attention_bwd(q, k, v, out, dout, lse, B, H, S, strides, klen, 0.125, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2])
ops.attention_bwd(q, k, v, out, dout, lse, B, H, S, strides, klen, 0.125, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2], drop_mask=mask, drop_p=0.1)
torch.cuda.synchronize()
