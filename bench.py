#!/usr/bin/env python
"""bench.py -- SDXL-1024 PEA KD training throughput (BASELINE.json's metric) on N B200s of one node.

    python bench.py --gpus 1 --steps K --warmup W                      # this repo's CUDA hot path
    torchrun ... bench.py --gpus N --steps K --warmup W                # one rank per GPU, NCCL adapter-grad all-reduce
    python bench.py --impl reference --gpus N --steps K --warmup W     # the reference's algorithm on the host cores

A step = one KD training step (train_sdxl_zh.py:305-449 from latents onward): batched adapter forward, CFG-dropout
select, student UNet forward (70 fused attn2 blocks), teacher UNet forward (no_grad), fused KD loss + gradient, backward
to the adapter through the student (dQ / dContext only), adapter-grad all-reduce, AdamW on the adapter.
Workload (config.workload): BASELINE.json configs[1] -- random-init SDXL-shaped UNet shell x2, 1024 px (128x128 latent),
bf16, batch 8 per GPU, 77x1024 CLIP-shaped embeddings -> shipped 6.03 M adapter; synthetic data.

One JSON line on stdout (rank 0).  `value` = samples/s with inputs resident in HBM; `e2e` = the same with pinned-host
inputs copied H2D and the loss read back D2H inside every timed step; `roofline` = the fused attn2 forward+backward
kernels' algorithmic bytes / CUDA-event time over the timed region against MEASURED_PEAKS.json; `cpu_baseline` = the
oracle (CPU restatement of the reference path) timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "SDXL-1024 KD train samples/s"
UNIT = "samples/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--batch", type=int, default=8, help="per-GPU batch (config 2: 8)")
    ap.add_argument("--latent", type=int, default=128, help="latent side (128 = 1024 px)")
    ap.add_argument("--adapter-precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager step instead of the CUDA-graph replay")
    ap.add_argument("--overlap-teacher", type=int, default=1, help="1: teacher forward on its own stream, concurrent with the student forward (0: in line)")
    ap.add_argument("--timed-only", action="store_true", help="skip the eager per-kernel pass and the e2e loop (ncu launch-list runs)")
    ap.add_argument("--tiny", action="store_true", help="tiny shell (debug only; the JSON line says so)")
    return ap.parse_args()


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


# ------------------------------------------------------------------------------------------------- synthetic data
def synth_batch(B, latent, cfg, in_dim, seed, pinned=False):
    """seeded N(0,1) latents/noise/embeddings, zh = [0,1]*, 10 % CFG mask."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    Dc = cfg.cross_attention_dim
    b = {
        "noisy_latents": r(B, 4, latent, latent), "noise": r(B, 4, latent, latent),
        "timesteps": torch.randint(0, 1000, (B,), generator=g),
        "enc": r(B, 77, in_dim), "enc_uncond": r(B, 77, in_dim),
        "teacher_ctx": r(B, 77, Dc), "teacher_ctx_uncond": r(B, 77, Dc),
        "zh_or_not": torch.tensor([i % 2 for i in range(B)]),
        "prompt_mask": torch.rand(B, generator=g) < 0.1,
    }
    if cfg.addition_embed_type == "text_time":
        pooled = cfg.projection_class_embeddings_input_dim - 6 * cfg.addition_time_embed_dim
        b["teacher_pooled"] = r(B, pooled)
        b["time_ids"] = torch.tensor([[1024.0, 1024.0, 0.0, 0.0, 1024.0, 1024.0]] * B)
    return b


def host_stage(batch, act_dtype):
    """Host-side batch in the dtypes the step consumes, in pinned memory (what a DataLoader with pin_memory hands over)."""
    out = {}
    for k, v in batch.items():
        if k in ("enc", "enc_uncond"):
            v = v.float()
        elif v.is_floating_point():
            v = v.to(act_dtype)
        out[k] = v.contiguous().pin_memory()
    return out


def to_device(hbatch, dev):
    return {k: v.to(dev, non_blocking=True) for k, v in hbatch.items()}


def profile_in_graph(trainer, adapter, replays):
    """Capture trainer._forward_backward into a fresh CUDA graph with ops.TIMER in external-event mode, replay it
    ``replays`` times and return ({kernel: {launches, ms, bytes, flops}} summed over the replays, graph keep-alive)."""
    import torch
    from pea_diffusion_b200 import ops
    params = list(adapter.parameters())
    saved = [q.grad for q in params]                    # the trainer's own graph writes into these: keep them attached
    for q in params:
        q.grad = None
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    total = {}
    g = torch.cuda.CUDAGraph()
    try:
        ops.TIMER.reset(True, external=True)
        with torch.cuda.graph(g, stream=side):
            trainer._forward_backward(trainer._static_in)
        ops.TIMER.enabled = False
        for _ in range(replays):
            g.replay()
            torch.cuda.synchronize()
            for name, r in ops.TIMER.summary().items():
                t = total.setdefault(name, {"launches": 0, "ms": 0.0, "bytes": 0, "flops": 0})
                for k in t:
                    t[k] += r[k]
    finally:
        ops.TIMER.reset(False)
        for q, g_ in zip(params, saved):
            q.grad = g_
    return total, g


def event_pair_overhead_us(dev, in_graph, n=100):
    """Median elapsed time of a CUDA-event pair around the library's smallest launch (pea_cfg_select on 8 elements),
    taken the same way as the per-kernel figures (event-record nodes inside a replayed CUDA graph, or eager): what the
    bracketing itself adds to every per-launch time (record + launch latency + a ~2 us minimal kernel).  Reported next to
    the raw roofline numbers; never subtracted from `achieved` / `frac`."""
    import torch
    from pea_diffusion_b200 import ops
    m = torch.zeros(1, dtype=torch.uint8, device=dev)
    a = torch.zeros(1, 8, dtype=torch.bfloat16, device=dev)
    out = ops.cfg_select(m, a, a)   # warm-up (and the output buffer the captured launches reuse)
    pairs = []

    def body():
        for _ in range(n):
            e0 = torch.cuda.Event(enable_timing=True, external=in_graph)
            e1 = torch.cuda.Event(enable_timing=True, external=in_graph)
            e0.record()
            ops.cfg_select(m, a, a)
            e1.record()
            pairs.append((e0, e1))

    if in_graph:
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            body()
        g.replay()
        g.replay()
    else:
        body()
    torch.cuda.synchronize()
    t = sorted(e0.elapsed_time(e1) for e0, e1 in pairs)
    del out
    return t[len(t) // 2] * 1e3


# ------------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200",
                                       "-i", str(index)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        rows = [ln.strip().split(", ") for ln in open(self.f.name) if ln.strip()]
        os.unlink(self.f.name)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2]))
            except Exception:
                continue
            for n, v in zip(names, r[4:8]):
                if v.strip().lower() == "active":
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------------- reference arm
def cpu_reference_sample(steps: int, warmup: int, latent: int, batch: int, tiny: bool = False):
    """The reference's algorithm for the path (oracle/pea_oracle.py: stock attention processors, PyTorch adapter and
    loss, autograd INCLUDING the discarded UNet weight gradients, as the reference runs it) on the host cores.
    Returns (seconds per step, description)."""
    from oracle import pea_oracle as O
    from pea_diffusion_b200.shell.unet import UNet2DConditionShell, UNetConfig
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    cfg = UNetConfig.tiny("sdxl") if tiny else UNetConfig.sdxl()
    in_dim, hidden = (64, 64) if tiny else (1024, 1024)
    torch.manual_seed(0)
    student = UNet2DConditionShell(cfg)
    teacher = student  # same weights as the reference's teacher (loaded from the same checkpoint); shared to bound host RAM
    pooled = cfg.projection_class_embeddings_input_dim - 6 * cfg.addition_time_embed_dim
    sd = {k: v.requires_grad_(True) for k, v in O.adapter_init(in_dim, pooled, hidden, cfg.cross_attention_dim, seed=0).items()}
    opt = torch.optim.AdamW(list(sd.values()), lr=1e-5, weight_decay=0.0)
    b = synth_batch(batch, latent, cfg, in_dim, 1234)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        for p in student.parameters():
            p.grad = None
        loss, *_ = O.kd_train_step(student, teacher, sd, b, variant="sdxl")
        loss.backward()
        opt.step()
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return sum(times) / len(times), cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    latent, batch = (16, 2) if args.tiny else (32, 1)
    sec, cores = cpu_reference_sample(args.steps, args.warmup, latent, batch, args.tiny)
    # scale to the metric's unit: a 1024-px sample has (128/latent)^2 times the latent tokens of the timed sample
    scale = (args.latent / latent) ** 2
    value = batch / sec / scale
    sample = (f"oracle KD step (student fwd+bwd incl. discarded UNet weight grads, teacher fwd, PyTorch adapter+loss, AdamW), "
              f"fp32, SDXL shell, batch {batch}, {latent}x{latent} latent; samples/s divided by {scale:.0f} (token ratio to "
              f"{args.latent}x{args.latent}) -- linear-in-tokens scaling flatters the CPU (self-attention is quadratic)")
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic",
            "config": {"workload": "train_sdxl_zh.py KD step, random-init SDXL UNet shell x2, 1024px, batch 8/GPU (timed on a bounded CPU sample)",
                       "l2": "working set >> L2"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


# ------------------------------------------------------------------------------------------------- our arm
_REAL_STDOUT = None


def _guard_stdout():
    """stdout must carry exactly ONE JSON line.  Libraries print there too (NCCL's version banner, warnings from child
    threads): point fd 1 at stderr for the whole run and keep a private handle on the real stdout for the JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict) -> None:
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    args = parse()
    _guard_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback); use --impl reference for the host arm"
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # stdout carries exactly ONE JSON line: NCCL's banner / debug output (stdout by default) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    from pea_diffusion_b200 import _lib, adapter as A, attn2, ops
    from pea_diffusion_b200.shell.unet import UNet2DConditionShell, UNetConfig
    from pea_diffusion_b200.train import KDTrainer
    import copy

    act = torch.bfloat16
    cfg = UNetConfig.tiny("sdxl") if args.tiny else UNetConfig.sdxl()
    in_dim, hidden = (64, 64) if args.tiny else (1024, 1024)
    latent = 16 if args.tiny else args.latent
    pooled = cfg.projection_class_embeddings_input_dim - 6 * cfg.addition_time_embed_dim
    torch.manual_seed(0)  # identical weights on every rank
    with torch.device(dev):
        student = UNet2DConditionShell(cfg)
    student = student.to(act)
    teacher = copy.deepcopy(student)
    attn2.install(student, freeze=True)
    attn2.install(teacher, freeze=True)
    adapter = A.MLP(in_dim, pooled, hidden, cfg.cross_attention_dim, False, precision=args.adapter_precision).to(dev)
    trainer = KDTrainer(student, teacher, adapter, "sdxl", lr=1e-5, cuda_graph=not args.no_graph, overlap_teacher=bool(args.overlap_teacher))

    B = args.batch
    hbatch = host_stage(synth_batch(B, latent, cfg, in_dim, 1234 + rank), act)
    dbatch = to_device(hbatch, dev)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()

    def timed(fn, steps):
        barrier(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        torch.cuda.synchronize(); barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms) / steps

    try:
        for _ in range(max(args.warmup, 3)):
            trainer.step(dbatch)
        torch.cuda.synchronize()
    except Exception as exc:
        if not args.overlap_teacher:
            raise
        # the two-stream schedule is an optimisation, not part of the result: fall back to the in-line teacher forward
        sys.stderr.write("bench: teacher/student stream overlap failed (%s); re-running without it\n" % str(exc).splitlines()[0][:200])
        args.overlap_teacher = 0
        trainer.close()
        for q in adapter.parameters():
            q.grad = None
        trainer = KDTrainer(student, teacher, adapter, "sdxl", lr=1e-5, cuda_graph=not args.no_graph, overlap_teacher=False)
        for _ in range(max(args.warmup, 3)):
            trainer.step(dbatch)
        torch.cuda.synchronize()

    # ---- value: inputs resident in HBM ------------------------------------------------------------------------
    clocks = ClockSampler(local) if rank == 0 else None
    graph_mode = not args.no_graph
    n0 = _lib.launch_count()
    ops.TIMER.reset(not graph_mode)
    torch.cuda.nvtx.range_push("pea_timed")
    ms_step = timed(lambda: trainer.step(dbatch), args.steps)
    torch.cuda.nvtx.range_pop()
    ops.TIMER.enabled = False
    launches = _lib.launch_count() - n0
    prof_steps = args.steps
    if graph_mode and not args.timed_only:
        # A graph replay re-launches the kernels the capture recorded (launch_count only sees the capture), and CUDA
        # events cannot bracket nodes inside a graph: run the SAME step eagerly for the per-kernel roofline figures and
        # for the launch count.
        prof_steps = min(args.steps, 3)
        overlap_stream, trainer.teacher_stream = trainer.teacher_stream, None   # per-kernel times: one kernel at a time
        n0 = _lib.launch_count()
        ops.TIMER.reset(True)
        torch.cuda.nvtx.range_push("pea_profiled_eager")
        params = list(adapter.parameters())
        static_grads = [q.grad for q in params]          # the graph writes into these tensors: keep them attached
        for _ in range(prof_steps):
            for q in params:
                q.grad = None
            trainer._forward_backward(dbatch)
        torch.cuda.synchronize()
        for q, g_ in zip(params, static_grads):
            q.grad = g_
        torch.cuda.nvtx.range_pop()
        ops.TIMER.enabled = False
        launches = (_lib.launch_count() - n0) * args.steps // prof_steps
    ksum = ops.TIMER.summary()
    ops.TIMER.reset(False)
    timing_note = ("CUDA events around every launch of %d eager steps run right after the graph-replayed timed region"
                   % prof_steps) if graph_mode else "CUDA events around every launch inside the timed steps"
    if graph_mode and not args.timed_only:
        # Per-kernel times INSIDE a replayed step: capture the same forward+backward once more with external events
        # (event-record nodes) around every launch of the library, replay it, read the events after each replay.  The
        # eager figures above carry ~5-10 us of host launch latency per bracket; these do not.
        try:
            g_steps = min(args.steps, 5)
            ksum_g, _keep = profile_in_graph(trainer, adapter, g_steps)
            if ksum_g.get("attn2_fwd", {}).get("launches", 0) > 0:
                ksum, prof_steps = ksum_g, g_steps
                timing_note = ("CUDA events recorded as event-record nodes around every launch inside %d replays of a CUDA graph of "
                               "the same forward+backward, captured and replayed right after the timed region" % g_steps) + \
                              ("; captured WITHOUT the teacher/student stream overlap of the timed region, so that each kernel has the GPU "
                               "to itself while it is timed" if args.overlap_teacher else "")
            del _keep
        except Exception as exc:  # keep the eager figures; say why
            timing_note += " (in-graph event timing failed: %s)" % (str(exc).splitlines()[0][:160] if str(exc) else type(exc).__name__)
        trainer.teacher_stream = overlap_stream
        torch.cuda.empty_cache()
    try:
        ev_overhead_us = event_pair_overhead_us(dev, in_graph=timing_note.startswith("CUDA events recorded as event-record nodes"))
    except Exception:
        ev_overhead_us = None

    # ---- e2e: pinned host inputs -> H2D every step, loss read back D2H every step -------------------------------
    host_loss = torch.empty(4, dtype=torch.float32).pin_memory()

    def e2e_step():
        out = trainer.step(to_device(hbatch, dev))
        host_loss.copy_(out, non_blocking=False)  # D2H read of the step's result (blocks like .item() would)

    if args.timed_only:
        ms_e2e = ms_step
    else:
        e2e_step()
        ms_e2e = timed(e2e_step, args.steps)
    clk = clocks.stop() if clocks else None
    h2d = sum(v.numel() * v.element_size() for v in hbatch.values())

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk, pk_kind = peaks()
    at = {k: ksum.get(k, {"launches": 0, "ms": 0.0, "bytes": 0, "flops": 0}) for k in ("attn2_fwd", "attn2_bwd", "kd_mse", "gemm")}
    a_ms = at["attn2_fwd"]["ms"] + at["attn2_bwd"]["ms"]
    a_bytes = at["attn2_fwd"]["bytes"] + at["attn2_bwd"]["bytes"]
    achieved = a_bytes / (a_ms * 1e-3) / 1e9 if a_ms > 0 else 0.0
    # DRAM traffic per launch (read + write) of the same launch mix, from the committed ncu --set full capture of the two
    # shapes a step launches (profiles/r1_attn2_traffic.json); null on other workloads
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "r1_attn2_traffic.json")))
        if not args.tiny and args.latent == 128 and args.batch == 8:
            traffic, traffic_src = tj["traffic_bytes_per_launch_step_mix"], "profiles/r1_attn2_traffic.json (bytes per launch, step mix; algorithmic: %d)" % tj["algorithmic_bytes_per_launch_step_mix"]
    except Exception:
        pass
    per_kernel = {k: {"launches_per_step": v["launches"] / prof_steps, "ms_per_step": v["ms"] / prof_steps,
                      "GBps": (v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0.0),
                      "TFLOPs": (v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else 0.0)} for k, v in at.items()}
    value = world * B / (ms_step * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
        "data": "synthetic",
        "config": {"workload": ("TINY DEBUG SHELL -- not the benchmark config" if args.tiny else
                                "train_sdxl_zh.py KD step, random-init SDXL UNet shell x2 (student+teacher), 1024px "
                                f"({latent}x{latent} latent), bf16, batch {B}/GPU, teacher+student forward, backward to the 6.03M adapter, AdamW"),
                   "global_batch": world * B, "parallelism": f"dp{world}", "adapter_precision": args.adapter_precision,
                   "step_execution": ("CUDA graph replay (fwd+bwd captured once) + eager all-reduce/AdamW" if graph_mode else "eager") +
                                     ("; teacher forward on a parallel stream / graph branch" if args.overlap_teacher else ""),
                   "l2": "no explicit flush: every step streams 10.3 GB of weights plus activations (>> 126 MB L2)"},
        "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 16,
                "ms_per_step": ms_e2e},
        "gpu_launches": launches,
        "clocks": clk,
        "roofline": {"bound": "hbm", "kernel": "attn2 fused forward + backward (all 70 layers x student fwd/bwd + teacher fwd)",
                     "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
                     "traffic": traffic, "traffic_source": traffic_src, "peak_source": pk_kind,
                     "event_pair_overhead_us": ev_overhead_us,
                     "frac_net_of_event_overhead": (a_bytes / (max(a_ms - ev_overhead_us * 1e-3 * (at["attn2_fwd"]["launches"] + at["attn2_bwd"]["launches"]), 1e-9) * 1e-3) / 1e9 / pk["hbm_gbs"]) if (a_ms > 0 and ev_overhead_us is not None) else None,
                     "timing": timing_note,
                     "per_kernel": per_kernel},
    }
    if not args.no_cpu_baseline and world == 1:
        try:
            cl, cb = (16, 2) if args.tiny else (32, 1)
            sec, cores = cpu_reference_sample(2, 1, cl, cb, args.tiny)
            scale = (latent / cl) ** 2
            line["cpu_baseline"] = {"value": cb / sec / scale, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"oracle KD step fp32, SDXL shell, batch {cb}, {cl}x{cl} latent, {sec:.2f} s/step; samples/s "
                                              f"divided by {scale:.0f} (token ratio to {latent}x{latent})"}
        except Exception as e:  # the GPU numbers stand on their own
            line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    emit(line)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
