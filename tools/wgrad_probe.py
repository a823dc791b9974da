import os, sys
sys.path.insert(0, "/root/repo")
sys.path.insert(0, os.getcwd())
import torch
from tools.gemm_bench import bench
f32 = torch.float32
for dbg in (0, 2):
    bench(f"vit-s fc2 wgrad dbg={dbg}", 384, 1536, 806912, a_major=1, b_major=1, out_dtype=f32, reps=10, tile_ns=((256, 16),), accumulate=True, dbg=dbg or 64)
    bench(f"bert fc2 wgrad dbg={dbg}", 768, 3072, 102400, a_major=1, b_major=1, out_dtype=f32, reps=10, tile_ns=((256, 16),), accumulate=True, dbg=dbg or 64)
