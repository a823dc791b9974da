"""A/B of GEMM tile configurations on one shape, interleaved rounds (robust to clock / power drift).
   python tools/gemm_ab.py M N K epi(0|1|3) [rounds]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simseg_b200 import ops
from simseg_b200._lib import EPI_BIAS_GELU, EPI_DGELU

M, N, K, epi = (int(x) for x in sys.argv[1:5])
rounds = int(sys.argv[5]) if len(sys.argv) > 5 else 6
bf = torch.bfloat16
A = torch.randn(M, K, device="cuda", dtype=bf)
B = torch.randn(N, K, device="cuda", dtype=bf)
kw = dict(M=M, N=N, K=K, epilogue=epi, out=torch.empty(M, N, device="cuda", dtype=bf))
if epi == EPI_BIAS_GELU:
    kw["bias"] = torch.randn(N, device="cuda")
    kw["aux"] = torch.empty(M, N, device="cuda", dtype=bf)
if epi == EPI_DGELU:
    kw["aux"] = torch.randn(M, N, device="cuda", dtype=bf)
    if not os.environ.get("AB_KEEP_GELU"):               # AB_KEEP_GELU=1: gelu(h) kept from the forward, no re-emit
        kw["aux2"] = torch.empty(M, N, device="cuda", dtype=bf)
    kw["col_sum"] = torch.zeros(N, device="cuda")
if epi == 0:
    kw["bias"] = torch.randn(N, device="cuda")
cfgs = [(tn, mode) for mode in (16, 32) for tn in (128, 192, 256)]
if epi in (EPI_BIAS_GELU, EPI_DGELU):
    # A/B of the activation epilogue: 16 warps (default) vs the round-1 8-warp version (reserved bit 64), CTA pairs, auto tile
    cfgs = [(0, 256), (0, 64), (256, 32 + 256), (256, 32 + 64), (128, 32 + 256), (128, 32 + 64)]
    if os.environ.get("AB_KEEP_GELU"):                   # early (default) vs late (bit 512) request of the pre-activation boxes
        cfgs = [(0, 0), (0, 512)]
res = {c: [] for c in cfgs}
for r in range(rounds + 1):
    for c in cfgs:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            ops.gemm(A, B, tile_n=c[0], _dbg=c[1], **kw)
        e1.record()
        torch.cuda.synchronize()
        if r > 0:
            res[c].append(e0.elapsed_time(e1) / 5)
for c in cfgs:
    v = sorted(res[c])
    tag = ("c1" if c[1] & 16 else "c2" if c[1] & 32 else "auto") + ("/8w" if c[1] & 64 else "") + ("/late-aux" if c[1] & 512 else "")
    print(f"{tag}/{c[0]}: min {v[0]:.3f} med {v[len(v) // 2]:.3f} ms  {2.0 * M * N * K / v[len(v) // 2] / 1e9:.0f} TF/s")
