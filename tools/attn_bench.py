"""Attention kernel microbenchmark (GPU box): fwd / bwd ms for the ViT-S and BERT shapes of the bench step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simseg_b200 import ops


def run(B, H, S, masked, reps=5):
    g = torch.Generator(device="cuda").manual_seed(1)
    D = H * 64
    qkv = torch.randn(B, S, 3, H, 64, device="cuda", generator=g).bfloat16()
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    strides = (S * 3 * D, 3 * D, 64)
    klen = torch.randint(8, S + 1, (B,), device="cuda", generator=g, dtype=torch.int32) if masked else None
    dout = torch.randn(B, S, D, device="cuda", generator=g).bfloat16()
    dqkv = torch.empty_like(qkv)
    out, lse = ops.attention_fwd(q, k, v, B, H, S, strides, klen, 0.125)

    def t(fn):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps
    res = []
    for impl in ("tc", "ts"):
        os.environ["SIMSEG_ATTN_FWD"] = impl
        fwd = t(lambda: ops.attention_fwd(q, k, v, B, H, S, strides, klen, 0.125, out=out, lse=lse))
        res.append(f"fwd[{impl}] {fwd:7.3f} ms ({4.0 * B * H * S * S * 64 / fwd / 1e9:6.1f} TF/s)")
    os.environ.pop("SIMSEG_ATTN_FWD", None)
    for impl in ("mma", "tc"):
        os.environ["SIMSEG_ATTN_BWD"] = impl
        ms = t(lambda: ops.attention_bwd(q, k, v, out, dout, lse, B, H, S, strides, klen, 0.125, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2]))
        res.append(f"bwd[{impl}] {ms:7.3f} ms ({10.0 * B * H * S * S * 64 / ms / 1e9:6.1f} TF/s)")
    os.environ.pop("SIMSEG_ATTN_BWD", None)
    print(f"B={B} H={H} S={S} masked={masked}: " + " | ".join(res), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1:
        run(int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), False, reps=1)
        sys.exit(0)
    run(4096, 6, 197, False)
    run(1024, 12, 197, False)
    run(4096, 12, 25, True)
    run(4096, 12, 77, True)
