"""attention_bwd_tc_kernel timed with pieces knocked out (SIMSEG_ATTN_DBG bits; results are WRONG, timing only).
Needs an instrumented library: SIMSEG_NVCC_DEFINES=-DSIMSEG_ATTN_KNOCKOUT python -m simseg_b200.build --force
(the run-time tests of these bits cost the production kernel 10 %: 1.92 -> 1.73 ms on the ViT-S layer once compiled out).
bits: 1 no MUFU in phase A | 2 no phase B at all (ld dP, math, dS stores) | 4 no P / dS smem stores | 8 no global stores in drains
      16 no gradient MMAs (dV, dK, dQ) | 32 no S / dP MMAs | 64 no tcgen05.ld of S / dP | 128 no phase A at all"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simseg_b200 import ops


def run(B, H, S, masked):
    g = torch.Generator(device="cuda").manual_seed(1)
    D = H * 64
    qkv = torch.randn(B, S, 3, H, 64, device="cuda", generator=g).bfloat16()
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    strides = (S * 3 * D, 3 * D, 64)
    klen = torch.randint(8, S + 1, (B,), device="cuda", generator=g, dtype=torch.int32) if masked else None
    dout = torch.randn(B, S, D, device="cuda", generator=g).bfloat16()
    dqkv = torch.empty_like(qkv)
    out, lse = ops.attention_fwd(q, k, v, B, H, S, strides, klen, 0.125)
    res = []
    for dbg in (0, 1, 2, 4, 8, 16, 32, 48, 64, 128, 130, 178, 255):
        os.environ["SIMSEG_ATTN_DBG"] = str(dbg)
        fn = lambda: ops.attention_bwd(q, k, v, out, dout, lse, B, H, S, strides, klen, 0.125, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2])
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            fn()
        e1.record()
        torch.cuda.synchronize()
        res.append(f"dbg={dbg}: {e0.elapsed_time(e1) / 5:.3f}")
    os.environ.pop("SIMSEG_ATTN_DBG", None)
    print(f"B={B} H={H} S={S}: " + " | ".join(res), flush=True)


run(4096, 6, 197, False)
run(4096, 12, 25, True)
