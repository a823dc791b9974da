"""Diagnostic for the tcgen05 GEMM: prints an error map per configuration (run on the GPU box)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simseg_b200 import ops


def run(M, N, K, a_major, b_major, tile_n, dtype=torch.bfloat16):
    g = torch.Generator(device="cuda").manual_seed(1)
    A = torch.randn(M, K, device="cuda", generator=g).to(dtype)
    B = torch.randn(N, K, device="cuda", generator=g).to(dtype)
    ref = A.float() @ B.float().T
    a_st = A.T.contiguous() if a_major else A
    b_st = B.T.contiguous() if b_major else B
    try:
        out = ops.gemm(a_st, b_st, M=M, N=N, K=K, a_major=a_major, b_major=b_major, out_dtype=torch.float32, tile_n=tile_n)
        torch.cuda.synchronize()
    except Exception as e:  # noqa
        print(f"M={M} N={N} K={K} maj=({a_major},{b_major}) bn={tile_n} {dtype}: EXC {e}")
        return
    err = (out - ref).abs()
    rel = err.max().item() / ref.abs().max().item()
    msg = f"M={M} N={N} K={K} maj=({a_major},{b_major}) bn={tile_n} {str(dtype)[6:]}: rel={rel:.2e}"
    if rel > 1e-3:
        bad = err > 1e-2 * ref.abs().max()
        rows = bad.any(1).nonzero().flatten()[:12].tolist()
        cols = bad.any(0).nonzero().flatten()[:12].tolist()
        msg += f" BAD frac={bad.float().mean().item():.3f} rows={rows} cols={cols} out[0,:4]={out[0,:4].tolist()} ref[0,:4]={ref[0,:4].tolist()}"
    print(msg, flush=True)


if __name__ == "__main__":
    for dtype in (torch.bfloat16, torch.float32):
        for maj in ((0, 0), (0, 1), (1, 1), (1, 0)):
            for bn in (128, 192, 256):
                run(256, 512, 256, maj[0], maj[1], bn, dtype)
    run(128, 128, 64, 0, 0, 128)
    run(4096, 1536, 384, 0, 0, 0)
    # quick throughput probe
    for (M, N, K) in ((807 * 1024, 1152, 384), (807 * 1024, 384, 1536), (201728, 3072, 768), (8192, 8192, 8192)):
        a = torch.randn(M, K, device="cuda").bfloat16(); b = torch.randn(N, K, device="cuda").bfloat16()
        out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
        for bn in (128, 192, 256):
            if N % bn and bn != 128:
                pass
            for _ in range(2):
                ops.gemm(a, b, M=M, N=N, K=K, out=out, tile_n=bn)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                ops.gemm(a, b, M=M, N=N, K=K, out=out, tile_n=bn)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            print(f"perf M={M} N={N} K={K} bn={bn}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.matmul(a, b.T, out=out)
        e0.record()
        for _ in range(5):
            torch.matmul(a, b.T, out=out)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"perf cuBLAS M={M} N={N} K={K}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.1f} TFLOP/s", flush=True)
        del a, b, out
