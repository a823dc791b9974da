"""Event trace of CTA 0 of the tcgen05 attention-backward kernel (simseg_debug_trace_*): per-phase clock deltas."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
from simseg_b200 import ops, _lib

B, H, S = (int(x) for x in sys.argv[1:4]) if len(sys.argv) > 3 else (2048, 6, 197)
g = torch.Generator(device="cuda").manual_seed(1)
D = H * 64
qkv = torch.randn(B, S, 3, H, 64, device="cuda", generator=g).bfloat16()
q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
strides = (S * 3 * D, 3 * D, 64)
dout = torch.randn(B, S, D, device="cuda", generator=g).bfloat16()
dqkv = torch.empty_like(qkv)
out, lse = ops.attention_fwd(q, k, v, B, H, S, strides, None, 0.125)
bw = lambda: ops.attention_bwd(q, k, v, out, dout, lse, B, H, S, strides, None, 0.125, dqkv[:, :, 0], dqkv[:, :, 1], dqkv[:, :, 2])
bw(); torch.cuda.synchronize()
lib = _lib.load()
lib.simseg_debug_trace_enable(1)
bw(); torch.cuda.synchronize()
NID = 24
buf = np.zeros(5 * 2 * NID, dtype=np.uint64)
lib.simseg_debug_trace_read(buf.ctypes.data_as(C.c_void_p), buf.size)
lib.simseg_debug_trace_enable(0)
names = {0: "?", 1: "loop top", 2: "wait p_ready", 3: "wait dkv_free", 4: "issue dV", 5: "issue S(n)", 6: "wait ds_ready", 7: "issue dP(n),dK,dQ",
         10: "loop back-edge", 11: "wait s_full", 12: "wait d_full + D/L read", 13: "phase A math", 14: "wait p_free",
         15: "P store+fence+arrive", 16: "wait dp_full", 17: "wait ds_free", 18: "phase B math+store+arrive", 19: "wait dkv_full",
         20: "drain dK/dV", 21: "wait dq_full", 22: "drain dQ"}
items = B * H / 148.0
for slot, label in enumerate(["MMA warp", "EW warp 2 (q2,c0)", "EW warp 3 (q3,c0)", "EW warp 6 (q2,c1)", "EW warp 10 (q2,c2)"]):
    acc = buf[slot * 2 * NID: slot * 2 * NID + NID].astype(np.int64)
    cnt = buf[slot * 2 * NID + NID: (slot + 1) * 2 * NID].astype(np.int64)
    tot = acc.sum()
    if tot == 0:
        continue
    print(f"== {label}: {tot} clocks total = {tot / items:.0f} per item")
    for i in range(NID):
        if cnt[i]:
            print(f"   {names.get(i, i):28s} n={cnt[i]:5d} mean {acc[i] / cnt[i]:7.0f} clk   {acc[i] / items:7.0f} clk/item  ({100 * acc[i] / tot:5.1f} %)")
