"""Patch-text similarity kernel microbenchmark (run on the GPU box): maps/s and achieved HBM GB/s at several batches.
   python tools/patch_sim_bench.py [C] [reps]"""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from simseg_b200 import ops


def run(B, N=196, C=171, E=512, reps=20):
    g = torch.Generator(device="cuda").manual_seed(3)
    bytes_per_map = N * E * 2 + N * C * 4 + N * 4
    nset = max(1, int(400e6 // (B * bytes_per_map)) + 1)            # cycle > 400 MB so nothing is served from L2
    ps = [torch.randn(B * N, E, device="cuda", generator=g).bfloat16() for _ in range(nset)]
    t = torch.nn.functional.normalize(torch.randn(C, E, device="cuda", generator=g), dim=-1).bfloat16()
    for i in range(nset):
        ops.patch_text_sim(ps[i], t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for i in range(nset):
            ops.patch_text_sim(ps[i], t)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / (reps * nset)
    gbs = B * bytes_per_map / ms / 1e6
    # device time per launch: the same nset launches replayed as one CUDA graph (no host issue path in the measurement)
    from simseg_b200.graph import replay_sequence
    gr = replay_sequence([(lambda x=x: ops.patch_text_sim(x, t)) for x in ps])
    gr.replay()
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        gr.replay()
    e1.record()
    torch.cuda.synchronize()
    msg = e0.elapsed_time(e1) / (reps * nset)
    print(json.dumps({"B": B, "N": N, "C": C, "ms": ms, "maps_per_s": B / ms * 1e3, "GBps": gbs, "nset": nset,
                      "ms_graph": msg, "GBps_graph": B * bytes_per_map / msg / 1e6}), flush=True)


if __name__ == "__main__":
    C = int(sys.argv[1]) if len(sys.argv) > 1 else 171
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
    for B in (64, 128, 256, 512, 4096):
        run(B, C=C, reps=reps if B < 4096 else max(2, reps // 4))
