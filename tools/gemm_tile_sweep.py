"""K-major bf16 GEMM time vs tile width for the BERT-side shapes of the sharded step (M = 25 tokens x per-GPU batch):
which BN wins once wave quantisation on 74 CTA pairs is counted.  python tools/gemm_tile_sweep.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simseg_b200 import ops


def t(fn, reps=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


shapes = [(768, 768), (768, 3072), (768, 2304), (2304, 768), (3072, 768), (384, 384), (384, 1536), (1152, 384), (1536, 384)]
for M in [int(a) for a in sys.argv[1:]] or [12800, 25600, 51200, 102400, 100864]:
    for N, K in shapes:
        a = torch.randn(M, K, device="cuda").bfloat16()
        b = torch.randn(N, K, device="cuda").bfloat16()
        bias = torch.randn(N, device="cuda")
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        res = []
        for bn in (0, 128, 192, 256):
            try:
                us = t(lambda: ops.gemm(a, b, M=M, N=N, K=K, bias=bias, out_dtype=torch.bfloat16, tile_n=bn, out=out))
                res.append(f"bn={bn:3d} {us:7.1f} us ({2.0 * M * N * K / us / 1e6:6.0f} TF/s)")
            except Exception as e:  # noqa: BLE001
                res.append(f"bn={bn:3d} n/a")
        print(f"M={M:6d} N={N:4d} K={K:4d}: " + " | ".join(res), flush=True)
