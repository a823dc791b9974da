#!/usr/bin/env python
"""Benchmark of the SimSeg hot path on B200 (contract: see README / DESIGN.md §Measurement).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun for N>1)
    python bench.py --impl reference --gpus N ...            # the reference's CPU path (oracle port) on host cores

Workload at every N (strong scaling): BASELINE.json configs[1] — ViT-S/16 224x224 + BERT-base contrastive
pre-training step (forward + backward + AdamW), GLOBAL batch 4096 image-text pairs, 25 text tokens, bf16
tensor-core operands, embeddings all-gathered across ranks before the InfoNCE logits.  Synthetic inputs,
seeded random-init weights.  One JSON line on stdout (rank 0).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "image-text pairs/sec (train)"
UNIT = "pairs/s"
BERT_DROPOUT = 0.1        # bert-base-uncased hidden_dropout_prob = attention_probs_dropout_prob, active in every timed train step
MODELS = {"vit-s": dict(yaml="simseg.vit-s.yaml", dim=384, heads=6, fwd_gf_img=9.20),
          "vit-b": dict(yaml="simseg.vit-b.yaml", dim=768, heads=12, fwd_gf_img=35.1)}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--model", default="vit-s", choices=list(MODELS))
    ap.add_argument("--global-batch", type=int, default=4096)
    ap.add_argument("--seq-len", type=int, default=25)
    ap.add_argument("--micro-batch", type=int, default=0, help="0 = whole per-rank batch in one pass")
    ap.add_argument("--cpu-sample", type=int, default=32, help="pairs in the CPU-baseline sample")
    ap.add_argument("--no-extras", action="store_true", help="skip roofline / patch-sim / cpu_baseline legs")
    ap.add_argument("--no-graph", action="store_true", help="issue every kernel from the host instead of replaying the CUDA graph of the step")
    return ap.parse_args()


def workload_name(a):
    return (f"ViT-{a.model[-1].upper()}/16 224x224 + BERT-base contrastive train step (fwd+bwd+AdamW), "
            f"global batch {a.global_batch}, {a.seq_len} text tokens, bf16")


# --------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock / power / throttle reasons sampled DURING the timed region: NVML in-process every 20 ms (nvidia-smi every
    200 ms if the NVML binding is unavailable)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.rows, self._stop, self._t = index, [], threading.Event(), None
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: honour CUDA_VISIBLE_DEVICES when it is a plain index list
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = index
            if vis:
                ids = [v.strip() for v in vis.split(",") if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self._h = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self._max = float(pynvml.nvmlDeviceGetMaxClockInfo(self._h, pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        sm = float(n.nvmlDeviceGetClockInfo(self._h, n.NVML_CLOCK_SM))
        pw = n.nvmlDeviceGetPowerUsage(self._h) / 1000.0
        r = n.nvmlDeviceGetCurrentClocksEventReasons(self._h) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") \
            else n.nvmlDeviceGetCurrentClocksThrottleReasons(self._h)
        flags = [bool(r & n.nvmlClocksThrottleReasonHwSlowdown), bool(r & n.nvmlClocksThrottleReasonHwThermalSlowdown),
                 bool(r & n.nvmlClocksThrottleReasonSwThermalSlowdown), bool(r & n.nvmlClocksThrottleReasonSwPowerCap)]
        self.rows.append([str(sm), str(self._max), str(pw)] + ["Active" if f else "Not Active" for f in flags])

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    o = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index)],
                                       capture_output=True, text=True, timeout=5).stdout.strip()
                    if o:
                        self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            self._stop.wait(0.02 if self._nvml is not None else 0.2)

    def __enter__(self):
        self._t = threading.Thread(target=self._run, daemon=True)
        self._t.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._t.join(timeout=6)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        reasons = [n for i, n in enumerate(self.NAMES)
                   if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons, "samples": len(sm),
                "power_w_max": max(float(r[2]) for r in self.rows if r[2].replace(".", "").isdigit()),
                "source": "nvml" if self._nvml is not None else "nvidia-smi"}


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return d["hbm_gbs"], d["bf16_tflops_sustained"], d["bf16_tflops"], "measured"
    return 6650.0, 1400.0, 1590.0, "fallback"


# --------------------------------------------------------------------------------------------- CPU arm
def cpu_pairs_per_s(a, pairs: int, steps: int = 1, warmup: int = 0):
    """The reference's path restated on the CPU (oracle port): fwd + bwd + AdamW on `pairs` image-text pairs."""
    import torch
    from oracle import simseg_oracle as O
    torch.set_num_threads(os.cpu_count())
    m = MODELS[a.model]
    sd = O.make_state_dict(m["dim"], m["heads"], seed=0)
    params = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in sd.items()}
    opt = torch.optim.AdamW([p for p in params.values() if p.requires_grad], lr=1e-4, betas=(0.9, 0.98), eps=1e-6,
                            weight_decay=0.001)
    batch = O.make_batch(pairs, a.seq_len, seed=1234)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        # train mode, as the reference runs it: BERT's hidden / attention-probability dropout (0.1) is part of the step
        loss, _, _ = O.clip_train_forward(params, batch, m["heads"], dropout=O.TorchDropout(BERT_DROPOUT, BERT_DROPOUT))
        loss.backward()
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    dt = sum(times) / len(times)
    return pairs / dt, dt, torch.get_num_threads()


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    v, dt, cores = cpu_pairs_per_s(a, a.cpu_sample, steps=a.steps, warmup=min(a.warmup, 1))
    sample = f"{a.cpu_sample}-pair micro-batch fwd+bwd+AdamW per step (a {a.global_batch}-pair CPU step would take hours)"
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": min(a.warmup, 1),
            "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference",
            "config": {"workload": workload_name(a), "global_batch": a.global_batch, "seq_len": a.seq_len,
                       "parallelism": "cpu"},
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------- our arm
def run_ours(a):
    import torch
    import torch.distributed as dist
    from simseg_b200 import ops
    from simseg_b200.synthetic import make_batch
    from simseg_b200.config import load_cfg
    from simseg_b200.pipeline import PIPELINE
    from simseg_b200.train import Trainer

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        import datetime
        if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
            os.environ["NCCL_DEBUG"] = "WARN"        # keep NCCL's version banner off stdout: one JSON line only
        # a collective mismatch must abort within minutes instead of hanging the box
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=240))
    assert a.global_batch % world == 0
    b = a.global_batch // world
    m = MODELS[a.model]
    cfg = load_cfg(m["yaml"], ["model.image_encoder.pretrained=False", "model.text_encoder.pretrained=False",
                               "transforms.input_size=224", f"data.batch_size={a.global_batch}"])
    torch.manual_seed(0)                            # same random-init weights on every rank (DDP broadcast not needed)
    model = PIPELINE["clip"](cfg).to(dev)
    use_graph = not a.no_graph and not a.micro_batch
    trainer = Trainer(model, cfg, micro_batch=a.micro_batch or None, capturable=use_graph)

    # synthetic per-rank shard, pinned on the host (two rotating host batches so every step copies fresh bytes)
    host = []
    for j in range(2):
        hb = make_batch(b, a.seq_len, seed=1234 + 17 * rank + 1000 * j)
        host.append({k: v.pin_memory() for k, v in hb.items()})
    h2d_bytes = sum(v.numel() * v.element_size() for v in host[0].values())
    copy_stream = torch.cuda.Stream()
    dev_bufs = [{k: torch.empty_like(v, device=dev) for k, v in host[0].items()} for _ in range(2)]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn(steps)
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return ms.item()

    # ---- dp_check (SURVEY §4 item 3 on hardware): every rank takes its slice of ONE fixed, seeded 64-pair global batch and
    # runs forward + backward + gradient all-reduce (no optimizer step).  The global-mean loss and the norms of the reduced
    # per-tower gradients must not depend on N: compare this block across the N = 1/2/4/8 lines.
    dp = dp_check(trainer, dev, world, rank, a.seq_len)

    # ---- the step: one CUDA-graph launch (forward, backward, all-reduces, AdamW, bf16 weight re-cast recorded once); the
    # eager path (every kernel issued from the host) is kept behind --no-graph and as the fallback if capture fails
    for k in dev_bufs[0]:
        dev_bufs[0][k].copy_(host[0][k])
    graphed, graph_note = None, "off (--no-graph / micro-batching)"
    launches_per_step = None
    if use_graph:
        try:
            graphed = trainer.capture(dev_bufs[0], warmup=max(a.warmup, 2))
            launches_per_step = graphed.launches_per_replay
            graph_note = "whole step replayed as one CUDA graph"
        except Exception as e:                               # noqa: BLE001 - report and fall back, never lose the line
            graphed, graph_note = None, f"capture failed, eager fallback: {repr(e)[:160]}"
            torch.cuda.synchronize()
    last = {}

    def do_step(batch, resident_inputs=False):
        if graphed is None:
            return trainer.step(batch)
        return graphed(None if resident_inputs else batch)    # None: the static input buffers already hold the batch

    # ---- resident-input loop (value): inputs already in HBM
    def resident(steps):
        for _ in range(steps):
            last["out"] = do_step(dev_bufs[0], resident_inputs=True)

    resident(a.warmup)
    ops.launch_count(reset=True)
    with ClockSampler(local) as cs:
        ms_total = timed(resident, a.steps)
    launches = ops.launch_count() if graphed is None else launches_per_step * a.steps
    ms_step = ms_total / a.steps
    value = a.global_batch / (ms_step / 1e3)

    # ---- end-to-end loop: pinned host batch -> H2D (prefetched on a copy stream into a staging buffer) -> step -> loss.item()
    # (a graph reads fixed addresses: the staged batch is copied device-to-device into the step's static input buffers)
    def e2e(steps):
        ready = [torch.cuda.Event(), torch.cuda.Event()]
        done = [None, None]

        def issue(i):
            s = i % 2
            with torch.cuda.stream(copy_stream):
                if done[s] is not None:
                    copy_stream.wait_event(done[s])           # the step that last read this buffer has finished
                for k in dev_bufs[s]:
                    dev_bufs[s][k].copy_(host[s][k], non_blocking=True)
                ready[s].record(copy_stream)
        issue(0)
        for i in range(steps):
            s = i % 2
            if i + 1 < steps:
                issue(i + 1)
            torch.cuda.current_stream().wait_event(ready[s])
            loss, _, _ = do_step(dev_bufs[s])
            done[s] = torch.cuda.Event()
            done[s].record()
            last["loss"] = loss.item()                        # D2H read of the step result
    e2e(1)
    ms_e2e = timed(e2e, a.steps) / a.steps
    e2e_value = a.global_batch / (ms_e2e / 1e3)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16",
            "data": "synthetic",
            "config": {"workload": workload_name(a), "global_batch": a.global_batch, "per_gpu_batch": b,
                       "seq_len": a.seq_len, "parallelism": f"dp{world}", "micro_batch": a.micro_batch or b,
                       "cuda_graph": graph_note,
                       "bert_dropout": {"hidden": model.text_encoder.model.hidden_dropout_prob,
                                        "attention_probs": model.text_encoder.model.attention_probs_dropout_prob,
                                        "mode": "model.train(): masks drawn in-kernel (Philox), fresh every replay"},
                       "l2": "per-step working set (>10 GB of activations) is far larger than the 126 MB L2"},
            "clocks": cs.summary(),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d_bytes * world, "d2h_bytes_per_step": 4 * world,
                    "ms_per_step": ms_e2e},
            "gpu_launches": int(launches),
            "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9,
            "loss": float(last.get("loss", float("nan"))),
            "dp_check": dp}

    if not a.no_extras:
        hbm, tf_sus, tf_burst, how = peaks()
        graphed = None                                               # the graph's memory pool goes back before the extra legs
        last.clear()
        torch.cuda.empty_cache()
        roof = gemm_roofline(trainer, dev_bufs[0], tf_sus, how)      # a training step: EVERY rank takes part in its collectives
        if rank == 0:
            line["roofline"] = roof
            line["train_flops"] = train_flops(a, m, ms_step, tf_sus)
            line["patch_sim"] = patch_sim_bench(hbm, how)
        del trainer, model, roof
        torch.cuda.empty_cache()
        if world == 8 and a.model == "vit-s" and a.global_batch == 4096:
            # BASELINE configs[2]: ViT-B/16 contrastive pretrain, global batch 8192 on 8 GPUs (b = 1024 per GPU)
            try:
                cfg3 = train_leg("vit-b", 8192, a.seq_len, world, rank, dev, steps=4, warmup=3, tf_sus=tf_sus, how=how)
            except Exception as e:
                cfg3 = {"error": repr(e)[:300]}
            if rank == 0:
                line["cfg3_vit_b_gb8192"] = cfg3
        if rank == 0 and world == 1:
            try:
                line["inference"] = inference_extras(hbm, tf_sus)
            except Exception as e:                       # never lose the headline line to an extra
                line["inference"] = {"error": repr(e)[:200]}
            try:
                line["cfg1_vit_s_b32_t77_c20"] = cfg1_forward(hbm, tf_sus)
            except Exception as e:
                line["cfg1_vit_s_b32_t77_c20"] = {"error": repr(e)[:200]}
    if world > 1:
        dist.barrier()
    if rank == 0:
        if world == 1 and not a.no_extras:
            v, dt, cores = cpu_pairs_per_s(a, a.cpu_sample, steps=4, warmup=1)      # ~10-15 s of host work
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{a.cpu_sample}-pair micro-batches fwd+bwd+AdamW of the same model, fp32: mean of 4 "
                                              f"steps after 1 warm-up, {dt:.1f} s per step"}
        print(json.dumps(line), flush=True)
    if world > 1:
        # The line is out.  Tearing down NCCL communicators that a captured CUDA graph still references blocked for minutes
        # (seen at N = 2: the run printed its line and then sat in destroy_process_group until the harness killed it), so
        # every rank synchronises once more and leaves without the teardown.
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)


def fwd_gflop_per_pair(m, T):
    return m["fwd_gf_img"] + T * 12 * (24 * 768 ** 2 + 4 * T * 768) / 1e9 + (2 * 196 * m["dim"] * 512 + 2 * T * 768 * 512) / 1e9


def train_flops(a, m, ms_step, tf_peak):
    """Whole-step algorithmic FLOPs per SURVEY.md §8d (train = 3x forward)."""
    fwd_pair = fwd_gflop_per_pair(m, a.seq_len)
    tf = 3 * fwd_pair * a.global_batch / 1e3 / (ms_step / 1e3)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    return {"gflop_per_pair_train": 3 * fwd_pair, "achieved_tflops_all_gpus": tf,
            "frac_of_sustained_bf16_per_gpu": tf / world / tf_peak}


def dp_check(trainer, dev, world, rank, T, pairs=64):
    import torch
    import torch.distributed as dist
    from simseg_b200.synthetic import make_batch
    gb = make_batch(pairs, T, seed=4242)                         # the same global batch on every rank, whatever N is
    b = pairs // world
    mine = {k: v[rank * b:(rank + 1) * b].to(dev) for k, v in gb.items()}
    # eval mode for this block only: BERT's dropout masks are drawn per rank-local row, so a train-mode loss is not comparable
    # across N (the timed steps below run in train mode, dropout on, as the reference trains)
    trainer.model.eval()
    try:
        loss, i2t, t2i = trainer.backward_only(mine)
    finally:
        trainer.model.train()
    stats = torch.stack([loss.float(), i2t.float(), t2i.float()])
    if world > 1:
        dist.all_reduce(stats)
        stats /= world
    torch.cuda.synchronize()
    out = {"pairs": pairs, "loss": stats[0].item(), "i2t_acc": stats[1].item(), "t2i_acc": stats[2].item(),
           "grad_norm": {k: f.flat.double().norm().item() for k, f in trainer.flat.items()},
           "note": "global-mean loss and all-reduced gradient norms of one fixed 64-pair batch: must agree across N"}
    trainer.zero_grad()
    return out


def train_leg(model_key, global_batch, T, world, rank, dev, steps, warmup, tf_sus, how):
    """An extra timed training leg (its own model + Trainer), same timing rules as the headline: W warm-up steps, K steps
    between barrier + synchronize, CUDA events, max over ranks."""
    import torch
    import torch.distributed as dist
    from simseg_b200.config import load_cfg
    from simseg_b200.pipeline import PIPELINE
    from simseg_b200.synthetic import make_batch
    from simseg_b200.train import Trainer
    m = MODELS[model_key]
    b = global_batch // world
    cfg = load_cfg(m["yaml"], ["model.image_encoder.pretrained=False", "model.text_encoder.pretrained=False",
                               "transforms.input_size=224", f"data.batch_size={global_batch}"])
    torch.manual_seed(0)
    model = PIPELINE["clip"](cfg).to(dev)
    trainer = Trainer(model, cfg)
    batch = {k: v.to(dev) for k, v in make_batch(b, T, seed=99 + rank).items()}
    torch.cuda.reset_peak_memory_stats()
    for _ in range(warmup):
        out = trainer.step(batch)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        out = trainer.step(batch)
    e1.record()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = ms.item()
    roof = gemm_roofline(trainer, batch, tf_sus, how)
    fwd = fwd_gflop_per_pair(m, T)
    tf = 3 * fwd * global_batch / 1e3 / (ms / 1e3)
    res = {"workload": f"ViT-{model_key[-1].upper()}/16 224x224 + BERT-base contrastive train step (fwd+bwd+AdamW), global batch "
                       f"{global_batch}, {T} text tokens, bf16, dp{world} (per-GPU batch {b})",
           "value": global_batch / (ms / 1e3), "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": warmup,
           "loss": out[0].item(), "peak_mem_gb": torch.cuda.max_memory_allocated() / 1e9,
           "train_flops": {"gflop_per_pair_train": 3 * fwd, "achieved_tflops_all_gpus": tf,
                           "frac_of_sustained_bf16_per_gpu": tf / world / tf_sus},
           "roofline": roof}
    del trainer, model, batch
    torch.cuda.empty_cache()
    return res


def cfg1_forward(hbm_peak, tf_peak):
    """BASELINE.json configs[0]: ViT-S/16 224x224 + 77-token text, batch 32, 20-class patch-text similarity map, one GPU,
    forward only — image tower -> projection -> (32,196,20) map + argmax; text tower -> (32,512); image-text logits (32,32).
    The reference's CPU path (oracle port, fp32, all host cores) is timed beside it on the same inputs."""
    import torch
    from oracle import simseg_oracle as O                         # cpu_baseline leg only
    from simseg_b200 import ops
    from simseg_b200.config import load_cfg
    from simseg_b200.pipeline import PIPELINE
    cfg = load_cfg("simseg.vit-s.yaml", ["model.image_encoder.pretrained=False", "model.text_encoder.pretrained=False",
                                          "transforms.input_size=224"])
    sd = O.make_state_dict(384, 6, seed=0)
    model = PIPELINE["clip"](cfg).to("cuda").eval()
    model.load_state_dict(sd)
    hb = [O.make_batch(32, 77, seed=500 + i) for i in range(4)]
    batches = [{k: v.cuda() for k, v in b.items()} for b in hb]
    text = torch.nn.functional.normalize(torch.randn(20, 512, generator=torch.Generator().manual_seed(5)), dim=-1)
    tdev = text.cuda()

    def fwd(image, input_ids, attention_mask):
        feat = model.forward_image_feature(image)
        img = model.forward_image_project(feat)
        proj = model.image_projection(feat)
        sim, am = ops.patch_text_sim(proj.contiguous(), tdev)
        txt = model.forward_text_project(model.forward_text_feature(input_ids, attention_mask), attention_mask)
        logits = (img @ txt.T) / 0.02                              # 32x32 glue, as mml_loss.py:73
        return sim, am, logits

    def time_it(call):
        for i in range(5):
            call(batches[i % 4])
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(20):
            call(batches[i % 4])
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / 20

    def eager(b):
        with torch.no_grad():
            return fwd(b["image"], b["input_ids"], b["attention_mask"])
    ms_eager = time_it(eager)
    # the same call recorded once and replayed as ONE CUDA-graph launch (simseg_b200.graph.GraphedCall): the copy of the
    # batch into the graph's static inputs is inside the timed region
    from simseg_b200.graph import GraphedCall
    gc = GraphedCall(fwd, batches[0]["image"], batches[0]["input_ids"], batches[0]["attention_mask"])
    ms = time_it(lambda b: gc(b["image"], b["input_ids"], b["attention_mask"]))
    # CPU port on the same batch (also a parity check of this very leg)
    torch.set_num_threads(os.cpu_count())
    b0 = hb[0]
    out_g = gc(batches[0]["image"], batches[0]["input_ids"], batches[0]["attention_mask"])
    torch.cuda.synchronize()
    sim_g, am_g, logits_g = [t.cpu() for t in out_g]
    sim_e, am_e, logits_e = [t.cpu() for t in eager(batches[0])]
    graph_equals_eager = bool(torch.equal(sim_g, sim_e) and torch.equal(am_g, am_e) and torch.equal(logits_g, logits_e))

    def cpu():
        with torch.no_grad():
            tok = O.vit_forward(sd, b0["image"], 6, O.IMG_PREFIX)
            rsim, ram = O.patch_text_sim(O.simple_projection(tok[:, 1:], sd["image_projection.linear.weight"]), text)
            ie = O.image_embed(tok, sd["image_projection.linear.weight"], 5)
            te = O.text_embed(O.bert_forward(sd, b0["input_ids"], b0["attention_mask"], 12, O.TXT_PREFIX),
                              sd["text_projection.linear.weight"], b0["attention_mask"], 1)
            return rsim, ram, ie @ te.T / 0.02
    cpu()
    t0 = time.perf_counter()
    rsim, ram, rlog = cpu()
    dt = time.perf_counter() - t0
    top2 = rsim.topk(2, -1)[0]
    safe = (top2[..., 0] - top2[..., 1]) > 2e-2
    gf = 32 * (9.20 + 77 * 12 * (24 * 768 ** 2 + 4 * 77 * 768) / 1e9 + (2 * 2 * 196 * 384 * 512 + 2 * 77 * 768 * 512 + 2 * 196 * 512 * 20) / 1e9)
    return {"workload": "ViT-S/16 224x224 + BERT-base 77 tokens, batch 32, forward only: 196x20 patch-text map + argmax, (32,512) "
                        "embeddings, 32x32 logits", "ms_per_batch": ms, "maps_per_s": 32 / (ms / 1e3), "pairs_per_s": 32 / (ms / 1e3),
            "launch": "one CUDA-graph replay per batch (%d library kernels), inputs copied into the graph's static buffers inside "
                      "the timed region" % gc.launches_per_replay,
            "ms_per_batch_eager": ms_eager, "graph_equals_eager_bitwise": graph_equals_eager,
            "achieved_tflops": gf / ms, "frac_of_sustained_bf16": gf / ms / tf_peak,
            "cpu_baseline": {"ms_per_batch": dt * 1e3, "maps_per_s": 32 / dt, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": "the same 32-pair batch, fp32, second of two runs"},
            "parity_vs_cpu_port": {"sim_max_abs_diff": (sim_g - rsim).abs().max().item(),
                                   "logits_max_abs_diff": (logits_g - rlog).abs().max().item(),
                                   "argmax_equal_where_margin_gt_2e-2": bool(torch.equal(am_g.long()[safe], ram[safe])),
                                   "argmax_mismatches_total": int((am_g.long() != ram).sum())},
            "note": "eager = ~600 kernel launches issued one by one through the C ABI (host-bound); the graph replay removes the "
                    "host from the path"}


def gemm_roofline(trainer, batch, tf_peak, how):
    """Dominant kernel = the tcgen05 GEMM engine: algorithmic FLOPs of every GEMM launch of one step divided by the
    summed CUDA-event durations of those launches (events on the launching stream, separate pass after the timed run)."""
    import torch
    from simseg_b200 import ops
    rec = []
    orig = ops.gemm

    def timed_gemm(a, b, **kw):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = orig(a, b, **kw)
        e1.record()
        rec.append((2.0 * kw["M"] * kw["N"] * kw["K"], e0, e1))
        return out
    # one untimed eager step first: the caching allocator is empty here (the CUDA graph's pool was just handed back), and a
    # cudaMalloc between e0.record() and the launch would idle the GPU inside the bracket
    trainer.step(batch)
    torch.cuda.synchronize()
    ops.gemm = timed_gemm
    try:
        trainer.step(batch)
        torch.cuda.synchronize()
    finally:
        ops.gemm = orig
    fl = sum(r[0] for r in rec)
    ms = sum(r[1].elapsed_time(r[2]) for r in rec)
    ach = fl / (ms / 1e3) / 1e12
    # DRAM bytes per launch (ncu dram__bytes_read.sum + dram__bytes_write.sum averaged over the GEMM launches of one step) come
    # from the committed capture of the SAME workload; other shapes / shards report null
    traffic, traffic_src = None, None
    try:
        with open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "profiles", "r02_gemm_traffic.json")) as f:
            t = json.load(f)
        M_img = batch["image"].shape[0]
        if t["launches"] == len(rec) and t["workload"] == "vit-s gb4096 T25 n1" and M_img == 4096 and abs(fl - 155.9e12) < 2e12:
            traffic, traffic_src = t["dram_bytes_per_launch"], t["source"]
    except (OSError, KeyError, ValueError):
        pass
    return {"kernel": "simseg::gemm_kernel (tcgen05+TMA)", "bound": "tensor", "achieved": ach, "peak": tf_peak, "unit": "TFLOP/s",
            "frac": ach / tf_peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": f"{how} sustained bf16",
            "launches_per_step": len(rec), "gemm_ms_per_step": ms, "flops_per_launch": fl / max(len(rec), 1)}


def _patch_sim_time(ps, t, reps):
    import torch
    from simseg_b200 import ops
    for x in ps:
        ops.patch_text_sim(x, t)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        for x in ps:
            ops.patch_text_sim(x, t)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / (reps * len(ps))


def patch_sim_bench(hbm_peak, how):
    """Dense patch-text similarity map (bf16 patch embeddings in HBM -> fp32 map + argmax), 196 patches x 171 classes.
    Two points: BASELINE.json configs[3] (batch 64: 21 MB per launch, latency-bound) and 4096 maps per launch (1.37 GB,
    the kernel's HBM-bound regime — the roofline figure).  Inputs are cycled / larger than L2 so no launch is L2-served."""
    import torch
    N, C, E = 196, 171, 512
    g = torch.Generator(device="cuda").manual_seed(3)
    t = torch.nn.functional.normalize(torch.randn(C, E, device="cuda", generator=g), dim=-1).bfloat16()
    bytes_per_map = N * E * 2 + N * C * 4 + N * 4
    # batch 64 (BASELINE config): 16 distinct input sets (>= 370 MB) are cycled
    ps64 = [torch.randn(64, N, E, device="cuda", generator=g).bfloat16() for _ in range(16)]
    ms64 = _patch_sim_time(ps64, t, 10)
    # the same 16 launches recorded into one CUDA graph: device time per launch without the host's issue path
    from simseg_b200 import ops
    from simseg_b200.graph import replay_sequence
    g64 = replay_sequence([(lambda x=x: ops.patch_text_sim(x, t)) for x in ps64])
    for _ in range(3):
        g64.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        g64.replay()
    e1.record()
    torch.cuda.synchronize()
    ms64g = e0.elapsed_time(e1) / (10 * len(ps64))
    del g64, ps64
    # batch 4096: one 0.8 GB input, 0.55 GB output per launch
    big = [torch.randn(4096, N, E, device="cuda", generator=g).bfloat16()]
    ms4k = _patch_sim_time(big, t, 5)
    del big
    torch.cuda.empty_cache()
    ach = 4096 * bytes_per_map / (ms4k / 1e3) / 1e9
    ach64 = 64 * bytes_per_map / (ms64 / 1e3) / 1e9
    return {"workload": "ViT-B seg inference map: 196 patches x 171 classes per map, bf16 in, fp32 map + argmax",
            "value": 4096 / (ms4k / 1e3), "unit": "maps/s", "batch": 4096, "ms_per_batch": ms4k,
            "batch64": {"value": 64 / (ms64g / 1e3), "unit": "maps/s", "ms_per_batch": ms64g,
                        "achieved_gbs": 64 * bytes_per_map / (ms64g / 1e3) / 1e9,
                        "frac": 64 * bytes_per_map / (ms64g / 1e3) / 1e9 / hbm_peak,
                        "timing": "16 launches over 16 distinct input sets replayed as one CUDA graph (device time per launch)",
                        "host_launched": {"ms_per_batch": ms64, "achieved_gbs": ach64,
                                          "note": "one C-ABI call per launch from Python: the host's issue path (ctypes + two "
                                                  "tensor-map encodes) is longer than the kernel"},
                        "note": "BASELINE configs[3] batch: 21 MB per launch"},
            "roofline": {"bound": "hbm", "achieved": ach, "peak": hbm_peak, "unit": "GB/s", "frac": ach / hbm_peak,
                         "traffic": 1382.3e6, "traffic_source": "ncu --set full, dram__bytes_read+write per launch at batch 4096 "
                                                               "(profiles/r01_ncu_patch_sim_v3.txt); algorithmic 1374.4e6",
                         "peak_source": how, "algorithmic_bytes_per_map": bytes_per_map}}


def inference_extras(hbm_peak, tf_peak):
    """BASELINE.json configs[3] and configs[4] on one GPU (forward only, synthetic inputs, seeded random-init weights):
    ViT-B/16 seg inference (batch 64 -> 196 x 171 map per image, encoder + projection + map) and the 5k x 25k all-pairs
    retrieval ranks (fused tensor-core path)."""
    import torch
    from simseg_b200 import ops
    from simseg_b200.config import load_cfg
    from simseg_b200.pipeline import PIPELINE
    out = {}

    def t_ms(fn, reps, warm=3):
        for _ in range(warm):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    g = torch.Generator(device="cuda").manual_seed(7)
    cfg = load_cfg("simseg.vit-b.yaml", ["model.image_encoder.pretrained=False", "model.text_encoder.pretrained=False",
                                          "transforms.input_size=224"])
    torch.manual_seed(0)
    model = PIPELINE["clip"](cfg).to("cuda").eval()
    imgs = [torch.randn(64, 3, 224, 224, device="cuda", generator=g) for _ in range(4)]      # 154 MB of inputs cycled
    text = torch.nn.functional.normalize(torch.randn(171, 512, device="cuda", generator=g), dim=-1).bfloat16()
    state = {"i": 0}

    def seg_step():
        with torch.no_grad():
            x = imgs[state["i"] % 4]
            state["i"] += 1
            feat = model.forward_image_feature(x)
            proj = model.image_projection(feat)
            ops.patch_text_sim(proj.contiguous(), text if proj.dtype == torch.bfloat16 else text.float())
    ms_eager = t_ms(seg_step, 10)
    from simseg_b200.graph import GraphedCall

    def seg_fwd(x):
        feat = model.forward_image_feature(x)
        proj = model.image_projection(feat)
        return ops.patch_text_sim(proj.contiguous(), text if proj.dtype == torch.bfloat16 else text.float())
    gseg = GraphedCall(seg_fwd, imgs[0])

    def seg_graph():
        gseg(imgs[state["i"] % 4])                                # device-to-device copy of the batch + one graph launch
        state["i"] += 1
    ms = t_ms(seg_graph, 20)
    fl = 64 * (35.1e9 + 2 * 196 * 768 * 512 + 2 * 196 * 512 * 171)
    out["seg_infer_vit_b_b64"] = {"workload": "ViT-B/16 224x224, batch 64, encoder + projection + 196x171 map", "value": 64 / (ms / 1e3),
                                  "unit": "maps/s", "ms_per_batch": ms, "achieved_tflops": fl / (ms / 1e3) / 1e12,
                                  "frac_of_sustained_bf16": fl / (ms / 1e3) / 1e12 / tf_peak,
                                  "launch": "one CUDA-graph replay per batch (%d library kernels)" % gseg.launches_per_replay,
                                  "ms_per_batch_eager": ms_eager}
    del gseg
    del model, imgs
    torch.cuda.empty_cache()
    left = torch.nn.functional.normalize(torch.randn(5000, 512, device="cuda", generator=g), dim=-1)
    right = torch.nn.functional.normalize(torch.randn(25000, 512, device="cuda", generator=g), dim=-1)
    lg = torch.arange(5000, device="cuda", dtype=torch.int64)
    rg = torch.arange(25000, device="cuda", dtype=torch.int64) // 5
    holder = {}

    def retr():                                      # both directions, as tools/retrieval_evaluation.py evaluates them
        holder["i2t"] = ops.retrieval_rank_fused(left, right, lg, rg)
        holder["t2i"] = ops.retrieval_rank_fused(right, left, rg, lg)
    ms = t_ms(retr, 10, warm=3)
    fl = 2 * 2 * 2.0 * 5000 * 25000 * 512 * 3        # 2 directions x (best-match pass ~ +5 % not counted, rank pass) x 3 split products
    by = 2 * (5000 + 25000) * 512 * (4 + 4)          # fp32 read + bf16 hi/lo written, per direction (scores never reach HBM)

    def retr_simt():
        s = ops.allpairs_sim(left, right)
        holder["simt"] = ops.retrieval_rank(s, lg, rg)
    ms_simt = t_ms(retr_simt, 3, warm=1)
    agree = (holder["simt"] == holder["i2t"]).float().mean().item()
    out["retrieval_5k_x_25k"] = {"workload": "5000 x 25000 all-pairs cosine + first-match ranks, BOTH directions, split-bf16 tcgen05 "
                                             "products (fp32-grade), scores consumed in the GEMM epilogue (never written)",
                                 "ms": ms, "ms_per_direction": ms / 2, "value": 2 * 5000 * 25000 / (ms / 1e3), "unit": "pairs scored/s",
                                 "achieved_tflops_bf16": fl / 2 / (ms / 1e3) / 1e12,
                                 "frac_of_sustained_bf16": fl / 2 / (ms / 1e3) / 1e12 / tf_peak,
                                 "hbm_bytes_algorithmic": by, "r1_i2t": (holder["i2t"] == 0).float().mean().item(),
                                 "simt_fp32_one_direction_ms": ms_simt, "rank_agreement_with_simt_fp32": agree,
                                 "note": "tensor-bound: 2 x 5000 x 25000 x 512 MACs x 3 products per direction = 384 GFLOP -> 0.27 ms at "
                                         "the sustained bf16 peak; round 1 (fp32 SIMT GEMM + 500 MB matrix): 3.5 ms per direction"}
    return out


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)


if __name__ == "__main__":
    main()
