"""GEMM-engine microbenchmark over the shapes of one training step (run on the GPU box).
   python tools/gemm_bench.py [vit-s|vit-b|bert|all] [reps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simseg_b200 import ops
from simseg_b200._lib import EPI_BIAS_GELU, EPI_BIAS_RESIDUAL, EPI_DGELU, EPI_NONE

dev = "cuda"


def t_ms(fn, reps):
    for _ in range(3 if reps > 1 else 0):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def bench(tag, M, N, K, a_major=0, b_major=0, epi=EPI_NONE, out_dtype=torch.bfloat16, reps=10, tile_ns=(0,), accumulate=False, dbg=0):
    bf = torch.bfloat16
    A = torch.randn((K, M) if a_major else (M, K), device=dev, dtype=bf)
    B = torch.randn((K, N) if b_major else (N, K), device=dev, dtype=bf)
    kw = dict(M=M, N=N, K=K, a_major=a_major, b_major=b_major, epilogue=epi, out_dtype=out_dtype)
    out = torch.empty((M, N), device=dev, dtype=out_dtype)
    kw["out"] = out
    if epi in (EPI_BIAS_GELU, EPI_BIAS_RESIDUAL):
        kw["bias"] = torch.randn(N, device=dev)
    if epi == EPI_BIAS_RESIDUAL:
        kw["residual"] = torch.randn(M, N, device=dev)
    if epi == EPI_BIAS_GELU:
        kw["aux"] = torch.empty(M, N, device=dev, dtype=bf)
    if epi == EPI_DGELU:
        kw["aux"] = torch.randn(M, N, device=dev, dtype=bf)
        kw["col_sum"] = torch.zeros(N, device=dev)
        if out_dtype == bf:
            kw["aux2"] = torch.empty(M, N, device=dev, dtype=bf)
    if accumulate:
        kw["accumulate"] = True
    res = []
    for tn in tile_ns:
        # tn: tile_n, or (tile_n, mode) with mode 16 = single CTA, 32 = CTA pair; 0 = engine's own choice
        tile, mode = tn if isinstance(tn, tuple) else (tn, 0)
        kw["_dbg"] = dbg | mode
        try:
            ms = t_ms(lambda: ops.gemm(A, B, tile_n=tile, **kw), reps)
            res.append(f"{'auto' if not mode else ('c1' if mode == 16 else 'c2')}/{tile}: {ms:6.3f} ms {2.0 * M * N * K / ms / 1e9:6.0f}")
        except Exception as e:
            res.append(f"{mode}/{tile}: n/a")
    # cuBLAS reference point for the same math (library call, only as a yardstick)
    if not dbg:
        Ao, Bo = (A.t() if a_major else A), (B if b_major else B.t())
        ms = t_ms(lambda: torch.matmul(Ao, Bo), reps)
        res.append(f"cublas {ms:6.3f} ms {2.0 * M * N * K / ms / 1e9:6.0f}")
    print(f"{tag:34s} M={M:7d} N={N:5d} K={K:7d} | " + " | ".join(res), flush=True)


def tower(name, M, D, F, reps):
    f32 = torch.float32
    tn = (0, (128, 16), (192, 16), (256, 16), (128, 32), (192, 32), (256, 32))
    bench(f"{name} qkv fwd", M, 3 * D, D, reps=reps, tile_ns=tn)
    bench(f"{name} proj fwd", M, D, D, reps=reps, tile_ns=tn)
    bench(f"{name} fc1 fwd gelu+aux", M, F, D, epi=EPI_BIAS_GELU, reps=reps, tile_ns=tn)
    bench(f"{name} fc2 fwd", M, D, F, reps=reps, tile_ns=tn)
    bench(f"{name} fc2 dgrad dgelu+colsum", M, F, D, epi=EPI_DGELU, reps=reps, tile_ns=tn)
    bench(f"{name} fc1 dgrad", M, D, F, reps=reps, tile_ns=tn)
    bench(f"{name} qkv dgrad", M, D, 3 * D, reps=reps, tile_ns=tn)
    bench(f"{name} proj dgrad", M, D, D, reps=reps, tile_ns=tn)
    bench(f"{name} fc2 wgrad", D, F, M, a_major=1, b_major=1, out_dtype=f32, reps=reps, tile_ns=tn, accumulate=True)
    bench(f"{name} fc1 wgrad", F, D, M, a_major=1, b_major=1, out_dtype=f32, reps=reps, tile_ns=tn, accumulate=True)
    bench(f"{name} qkv wgrad", 3 * D, D, M, a_major=1, b_major=1, out_dtype=f32, reps=reps, tile_ns=tn, accumulate=True)
    bench(f"{name} proj wgrad", D, D, M, a_major=1, b_major=1, out_dtype=f32, reps=reps, tile_ns=tn, accumulate=True)


if __name__ == "__main__":
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which == "one":      # one M N K a_major b_major epi out(bf16|f32) tile_n [accumulate]
        M, N, K, am, bm, epi = (int(x) for x in sys.argv[2:8])
        od = torch.float32 if sys.argv[8] == "f32" else torch.bfloat16
        bench("one", M, N, K, a_major=am, b_major=bm, epi=epi, out_dtype=od, reps=2, tile_ns=(int(sys.argv[9]),),
              accumulate=len(sys.argv) > 10)
        sys.exit(0)
    if which == "prof":     # one launch per configuration, for `ncu --set full -k regex:gemm_kernel`
        f32 = torch.float32
        bench("vit-b fc1 dgrad (pair 256x256)", 201728, 768, 3072, reps=1, tile_ns=(0,))
        bench("vit-s fc2 fwd (pair 256x192)", 806912, 384, 1536, reps=1, tile_ns=(0,))
        bench("vit-s fc1 fwd gelu+aux", 806912, 1536, 384, epi=EPI_BIAS_GELU, reps=1, tile_ns=(0,))
        bench("vit-s fc2 dgrad dgelu+colsum", 806912, 1536, 384, epi=EPI_DGELU, reps=1, tile_ns=(0,))
        bench("vit-s fc2 wgrad (pair 256x384, dW^T)", 384, 1536, 806912, a_major=1, b_major=1, out_dtype=f32, reps=1, tile_ns=(0,), accumulate=True)
        sys.exit(0)
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    if which == "layouts":
        for (am, bm) in ((0, 0), (0, 1), (1, 0), (1, 1)):
            for (M, N, K) in ((8192, 8192, 4096), (4096, 4096, 65536)):
                bench(f"layout a_major={am} b_major={bm}", M, N, K, a_major=am, b_major=bm, reps=reps, tile_ns=(128, 256))
                for dbg in (1, 2, 4, 6):
                    bench(f"   dbg={dbg} ({ {1: 'no TMA', 2: 'no MMA', 4: '2-D boxes', 6: 'no MMA, 2-D boxes'}[dbg] })", M, N, K, a_major=am, b_major=bm, reps=reps,
                          tile_ns=(256,), dbg=dbg)
        sys.exit(0)
    if which in ("vit-s", "all"):
        tower("vit-s b=4096", 4096 * 197, 384, 1536, reps)
    if which in ("bert", "all"):
        tower("bert b=4096 T=25", 4096 * 25, 768, 3072, reps)
    if which in ("vit-b", "all"):
        tower("vit-b b=1024", 1024 * 197, 768, 3072, reps)
