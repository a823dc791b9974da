"""tcgen05 vs mma.sync forward attention at the segmentation-eval geometry (288 x 288 -> 325 tokens)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simseg_b200 import ops
for (B, H, S) in ((64, 6, 325), (64, 12, 325), (1024, 12, 325)):
    g = torch.Generator(device="cuda").manual_seed(1)
    D = H * 64
    qkv = torch.randn(B, S, 3, H, 64, device="cuda", generator=g).bfloat16()
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    strides = (S * 3 * D, 3 * D, 64)
    out = torch.empty(B, S, D, device="cuda", dtype=torch.bfloat16); lse = torch.empty(B, H, S, device="cuda")
    r = []
    for impl in ("mma", "tc"):
        os.environ["SIMSEG_ATTN_FWD"] = impl
        for _ in range(3):
            ops.attention_fwd(q, k, v, B, H, S, strides, None, 0.125, out=out, lse=lse)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ops.attention_fwd(q, k, v, B, H, S, strides, None, 0.125, out=out, lse=lse)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 20
        r.append(f"{impl} {ms * 1e3:8.1f} us ({4.0 * B * H * S * S * 64 / ms / 1e9:6.1f} TF/s)")
    print(f"B={B} H={H} S={S}: " + " | ".join(r), flush=True)
