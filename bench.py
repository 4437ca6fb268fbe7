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

One JSON line on stdout (rank 0):
  value          samples/s with inputs resident in HBM (weak scaling: batch 8 per GPU)
  e2e            the same with pinned-host inputs copied H2D and the loss read back D2H inside every timed step
  roofline       the fused attn2 forward+backward kernels: algorithmic bytes / CUDA-event time of
                 every launch inside a replayed CUDA graph of the step, against MEASURED_PEAKS.json; plus
                 .per_kernel        every kernel family of the library inside the step
                 .in_chain_marginal bracket-free: graph of 20 x [to_q GEMM -> attn2 -> to_out GEMM] minus the same graph without
                                    the attention launches, for the step's two shapes and both directions
                 .sweep             BASELINE.json configs[3]: the config-4 attn2 sweep (B=8; N 1024/4096/16384; C 640/1280;
                                    L 77, and 52 / 64 once), each launch alone, L2 flushed -- metric M2
  strong         configs[2]: the same step at GLOBAL batch 64 (64/W per rank, micro-batches of 8), samples/s
  infer          configs[4]: 30 forward-only UNet calls, fp16, UNet batch 8, constant context (K|V cache), vs stock processors
  dp_params_identical   bitwise equality of the adapter on every rank after the timed steps
  cpu_baseline   the oracle (CPU restatement of the reference path) timed on this box's host cores on a bounded sample
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
    ap.add_argument("--batch", type=int, default=8, help="per-GPU (micro-)batch (config 2: 8)")
    ap.add_argument("--global-batch", type=int, default=64, help="config 3's fixed global batch for the strong-scaling sub-run (0: skip it)")
    ap.add_argument("--latent", type=int, default=128, help="latent side (128 = 1024 px)")
    ap.add_argument("--dtype", default="bf16", choices=["bf16", "fp16"], help="UNet / activation dtype (fp16 = the reference's --precision 16, with loss scaling)")
    ap.add_argument("--adapter-precision", default=None, choices=["bf16", "fp16", "fp32"])
    ap.add_argument("--comm-dtype", default="fp32", choices=["fp32", "bf16"], help="payload of the adapter-gradient all-reduce")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager step instead of the CUDA-graph replay")
    ap.add_argument("--capture-comm", action="store_true", help="capture the NCCL all-reduce into the step's CUDA graph (experimental)")
    ap.add_argument("--overlap-teacher", type=int, default=1, help="1: teacher forward on its own stream, concurrent with the student forward (0: in line)")
    ap.add_argument("--timed-only", action="store_true", help="only the timed region (ncu launch-list runs)")
    ap.add_argument("--no-sweep", action="store_true")
    ap.add_argument("--no-infer", action="store_true")
    ap.add_argument("--tiny", action="store_true", help="tiny shell (debug only; the JSON line says so)")
    ap.add_argument("--workload", default="train", choices=["train", "infer"],
                    help="infer: BASELINE.json configs[4] alone (30 forward-only UNet calls, fp16, UNet batch 8) as its own JSON line")
    return ap.parse_args()


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))), "measured"
    except Exception:
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def workload_config(args, world):
    """The `config` object: identical for both arms (`--impl b200` and `--impl reference`)."""
    latent = 16 if args.tiny else args.latent
    return {"workload": ("TINY DEBUG SHELL -- not the benchmark config" if args.tiny else
                         f"train_sdxl_zh.py KD step, random-init SDXL UNet shell x2 (student+teacher), {8 * latent}px "
                         f"({latent}x{latent} latent), batch {args.batch}/GPU, 77x1024 embeddings -> 6.03M adapter, teacher+student "
                         "forward, backward to the adapter, AdamW"),
            "global_batch": world * args.batch, "parallelism": f"dp{world}",
            "l2": "no explicit flush in the step: every step streams 10.3 GB of weights plus activations (>> 126 MB L2); "
                  "roofline.sweep flushes L2 between iterations"}


# ------------------------------------------------------------------------------------------------- synthetic data
def synth_batch(B, latent, cfg, in_dim, seed):
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
    Imports nothing that binds libpea_b200.so.  Returns (seconds per step, cores)."""
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


def native_libs_mapped():
    """Names of this repo's shared objects mapped into the current process (the reference arm must list none)."""
    try:
        return sorted({ln.split()[-1] for ln in open("/proc/self/maps") if ROOT in ln and ln.rstrip().endswith(".so")})
    except Exception:
        return []


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    world = int(os.environ.get("WORLD_SIZE", str(args.gpus)))
    if args.tiny:
        latent, batch = 16, 2
        probe = None
    else:
        # bounded sample: one sample at a 64x64 latent (a quarter of the benchmark's tokens) when the whole --steps/--warmup
        # run then fits in ~4 minutes on this box's cores, else a 32x32 latent; decided from one timed 32x32 step
        sec32, _ = cpu_reference_sample(1, 1, 32, 1)
        probe = sec32
        latent, batch = (64, 1) if sec32 * 6.0 * (args.steps + args.warmup) <= 240.0 else (32, 1)
    sec, cores = cpu_reference_sample(args.steps, args.warmup, latent, batch, args.tiny)
    # scale to the metric's unit: a 1024-px sample has (128/latent)^2 times the latent tokens of the timed sample
    full = 16 if args.tiny else args.latent
    scale = (full / latent) ** 2
    value = batch / sec / scale
    sample = (f"oracle KD step (student fwd+bwd incl. discarded UNet weight grads, teacher fwd, PyTorch adapter+loss, AdamW), "
              f"fp32, SDXL shell, batch {batch}, {latent}x{latent} latent, {sec:.2f} s/step; samples/s divided by {scale:.0f} (token "
              f"ratio to {full}x{full}: an EXTRAPOLATION that flatters the CPU, self-attention being quadratic in tokens)" +
              (f"; 32x32 probe step {probe:.2f} s" if probe is not None else ""))
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "fp32", "data": "synthetic", "config": workload_config(args, world),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                             "note": "one host process whatever --gpus says: the CPU arm does not scale with N"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "native_libs_mapped": native_libs_mapped()}
    emit(line)


# ------------------------------------------------------------------------------------------------- our arm: helpers
_REAL_STDOUT = None


def _guard_stdout():
    """stdout must carry exactly ONE JSON line.  Libraries print there too (NCCL's version banner, warnings from child
    threads): point fd 1 at stderr for the whole run and keep a private handle on the real stdout for the JSON line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def log(msg: str) -> None:
    """Progress to stderr (a hung phase must be locatable from the driver's stderr tail)."""
    sys.stderr.write("[bench %7.1fs] %s\n" % (time.perf_counter() - _T0, msg))
    sys.stderr.flush()


_T0 = time.perf_counter()


def safe_destroy_process_group(grace_s: float = 60.0) -> None:
    """destroy_process_group() with a deadline: the results are already printed, a blocked NCCL teardown must not turn a
    finished run into a driver timeout."""
    import threading
    import torch.distributed as dist
    done = threading.Event()

    def bail():
        if not done.wait(grace_s):
            sys.stderr.write("[bench] process-group teardown still blocked after %.0f s; exiting\n" % grace_s)
            sys.stderr.flush()
            os._exit(0)

    threading.Thread(target=bail, daemon=True).start()
    dist.destroy_process_group()
    done.set()


def emit(line: dict) -> None:
    out = _REAL_STDOUT if _REAL_STDOUT is not None else sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def capture_fb(trainer, timer_external=False):
    """Capture trainer's forward+backward (no collectives) into a fresh CUDA graph -> (graph, library launches captured)."""
    from pea_diffusion_b200 import _lib, ops
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    world, trainer.reducer.world = trainer.reducer.world, 1      # measurement graphs carry no NCCL nodes
    try:
        with torch.cuda.stream(side):     # one eager pass on the capture stream first (as KDTrainer._capture does): per-pass caches
            trainer._forward_backward(trainer._static_in)   # and autograd state are then owned by this stream, not by an old graph
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        if timer_external:
            ops.TIMER.reset(True, external=True)
        with torch.cuda.graph(g, stream=side):
            trainer._forward_backward(trainer._static_in)
    finally:
        ops.TIMER.enabled = False
        trainer.reducer.world = world
    return g, _lib.launch_count() - n0


def profile_in_graph(trainer, replays):
    """Per-launch CUDA-event times INSIDE a replayed graph of the step (event-record nodes around every library launch).
    -> ({kernel: {launches, ms, bytes, flops}} summed over the replays, launches per step, graph keep-alive)."""
    from pea_diffusion_b200 import ops
    total = {}
    try:
        g, nlaunch = capture_fb(trainer, timer_external=True)
        for _ in range(replays):
            g.replay()
            torch.cuda.synchronize()
            for name, r in ops.TIMER.summary().items():
                t = total.setdefault(name, {"launches": 0, "ms": 0.0, "bytes": 0, "flops": 0})
                for k in t:
                    t[k] += r[k]
    finally:
        ops.TIMER.reset(False)
    return total, nlaunch, g


def time_replays(g, n):
    for _ in range(2):
        g.replay()
    torch.cuda.synchronize()
    ts = []
    for _ in range(n):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); e1.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def chain_marginal(dev, hbm_gbs, reps=20, replays=10, dtype=torch.bfloat16):
    """Bracket-free device time of one attention launch in the situation of a training step: a CUDA graph of `reps` x
    [to_q GEMM -> attn2 core -> to_out GEMM] (cuBLAS projections, as the default processor runs them; Q freshly written, so
    L2-resident) against the same graph without the attention launches; the difference / reps is what one launch adds --
    launch gap, prologue, fill and drain included, no event-record nodes.  Done for the two shapes and both directions an SDXL
    step launches, then weighted by the step's launch mix (120 + 20 forward, 60 + 10 backward)."""
    from pea_diffusion_b200 import ops
    out = {}
    total_ms, total_bytes = 0.0, 0
    for (N, C, H, n_fwd, n_bwd) in ((1024, 1280, 20, 120, 60), (4096, 640, 10, 20, 10)):
        B, L = 8, 77
        h = torch.randn(B * N, C, device=dev, dtype=dtype)
        wq = torch.randn(C, C, device=dev, dtype=dtype) * C ** -0.5
        wo = torch.randn(C, C, device=dev, dtype=dtype) * C ** -0.5
        kv = torch.randn(B, L, 2 * C, device=dev, dtype=dtype)
        q = torch.empty(B * N, C, device=dev, dtype=dtype)
        o2 = torch.empty(B * N, C, device=dev, dtype=dtype)
        do = torch.randn(B, N, C, device=dev, dtype=dtype)
        dkv = torch.zeros(B, L, 2 * C, device=dev, dtype=torch.float32)

        def body(mode):
            for _ in range(reps):
                torch.matmul(h, wq.t(), out=q)
                if mode == "fwd":
                    o = ops.attn2_fwd(q.view(B, N, C), kv, H, 0.125)
                elif mode == "bwd":
                    o, _ = ops.attn2_bwd(q.view(B, N, C), kv, do, H, 0.125, dkv=dkv)
                else:
                    o = do
                torch.matmul(o.view(B * N, C), wo.t(), out=o2)

        times = {}
        for mode in ("none", "fwd", "bwd"):
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                body(mode)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):
                body(mode)
            times[mode] = time_replays(g, replays) * 1e3     # us
            del g
        fb, bb = ops.attn2_bytes(B, N, L, C, 2)
        for mode, nbytes, count in (("fwd", fb, n_fwd), ("bwd", bb, n_bwd)):
            us = (times[mode] - times["none"]) / reps
            out[f"{mode}_N{N}_C{C}"] = {"us_per_launch": us, "GBps": nbytes / us / 1e3, "frac": nbytes / us / 1e3 / hbm_gbs,
                                         "launches_per_step": count}
            total_ms += us * count * 1e-3
            total_bytes += nbytes * count
    out["step_mix"] = {"attn2_ms_per_step": total_ms, "GBps": total_bytes / (total_ms * 1e-3) / 1e9,
                       "frac": total_bytes / (total_ms * 1e-3) / 1e9 / hbm_gbs}
    out["how"] = ("CUDA graph of %d x [cuBLAS to_q -> attn2 core -> cuBLAS to_out] minus the same graph without the attention launches, "
                  "median of %d replays; no event brackets; weighted by the launch mix of one SDXL step" % (reps, replays))
    return out


def event_pair_overhead_us(dev, n=100):
    """Median elapsed time of an event-record-node pair around the library's smallest launch inside a replayed graph: what
    the bracketing itself adds to every per-launch figure.  Reported beside the raw numbers; never subtracted from `frac`."""
    from pea_diffusion_b200 import ops
    m = torch.zeros(1, dtype=torch.uint8, device=dev)
    a = torch.zeros(1, 8, dtype=torch.bfloat16, device=dev)
    out = ops.cfg_select(m, a, a)
    pairs = []
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        for _ in range(n):
            e0 = torch.cuda.Event(enable_timing=True, external=True)
            e1 = torch.cuda.Event(enable_timing=True, external=True)
            e0.record()
            ops.cfg_select(m, a, a)
            e1.record()
            pairs.append((e0, e1))
    g.replay(); g.replay()
    torch.cuda.synchronize()
    t = sorted(e0.elapsed_time(e1) for e0, e1 in pairs)
    del out
    return t[len(t) // 2] * 1e3


def timed_isolated(fn, iters, scratch, warm=5):
    """Median CUDA-event time (us) of `fn` launched alone: L2 flushed by reading a 256 MB scratch (leaves clean lines), the
    host kept ahead of the device by a short spin kernel so no launch latency lands between the events."""
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        scratch.sum()
        torch.cuda._sleep(400000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        e1.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def attn2_sweep(dev, hbm_gbs, iters=20, dtype=torch.bfloat16):
    """BASELINE.json configs[3] / SURVEY.md section 8d config 4 (metric M2)."""
    from pea_diffusion_b200 import ops
    scratch = torch.empty(64 << 20, device=dev, dtype=torch.int32)
    rows = []
    B = 8
    for (C, H) in ((640, 10), (1280, 20)):
        for N in (1024, 4096, 16384):
            for L in ((77, 64, 52) if (N == 4096 and C == 640) else (77,)):
                q = torch.randn(B, N, C, device=dev, dtype=dtype)
                kv = torch.randn(B, L, 2 * C, device=dev, dtype=dtype)
                do = torch.randn(B, N, C, device=dev, dtype=dtype)
                dkv = torch.zeros(B, L, 2 * C, device=dev, dtype=torch.float32)
                d = C // H
                fb, bb = ops.attn2_bytes(B, N, L, C, 2)
                tf = timed_isolated(lambda: ops.attn2_fwd(q, kv, H, d ** -0.5), iters, scratch)
                tb = timed_isolated(lambda: ops.attn2_bwd(q, kv, do, H, d ** -0.5, dkv=dkv), iters, scratch)
                rows.append({"B": B, "N": N, "C": C, "H": H, "L": L, "fwd_us": round(tf, 2), "bwd_us": round(tb, 2),
                             "fwd_frac": round(fb / tf / 1e3 / hbm_gbs, 4), "bwd_frac": round(bb / tb / 1e3 / hbm_gbs, 4),
                             "fwdbwd_frac": round((fb + bb) / (tf + tb) / 1e3 / hbm_gbs, 4),
                             "fwd_TFLOPs": round(4 * B * N * L * C / tf / 1e6, 1), "bwd_TFLOPs": round(10 * B * N * L * C / tb / 1e6, 1)})
                del q, kv, do, dkv
    return rows


def gemm_sweep(dev, iters=20, dtype=torch.bfloat16):
    """The GEMM shapes the path launches (adapter chain at M = 2*8*77 rows, the all-layer K|V projection, the all-layer
    dContext) with cuBLAS timed beside each on the same operands -- context for the tensor-pipe figures."""
    from pea_diffusion_b200 import ops
    scratch = torch.empty(64 << 20, device=dev, dtype=torch.int32)
    shapes = [("adapter_fwd_1024x1024", 1232, 1024, 1024, True, 0), ("adapter_fwd_1280x1024", 1232, 1280, 1024, True, 0),
              ("adapter_fwd_fc_2048x1280", 1232, 2048, 1280, True, 0), ("adapter_dx_1024x1024", 1232, 1024, 1024, False, 0),
              ("adapter_dw_1024x1024", 1024, 1024, 1232, None, 0),
              ("kvproj_all_layers", 616, 166400, 2048, True, 0), ("dctx_all_layers", 616, 2048, 166400, False, 1)]
    rows = []
    for name, M, N, K, bk, splitk in shapes:
        if bk is None:        # dW = G^T X: both operands stored [rows = K, *]
            A = torch.randn(K, M, device=dev, dtype=dtype); Bm = torch.randn(K, N, device=dev, dtype=dtype)
            o = torch.empty(M, N, device=dev, dtype=torch.float32)
            fn = lambda: ops.gemm([A], [Bm], M, N, K, a_kmajor=False, b_kmajor=False, out0=o)
            lib = lambda: torch.matmul(A.t(), Bm)
        else:
            A = torch.randn(M, K, device=dev, dtype=dtype)
            Bm = torch.randn(N, K, device=dev, dtype=dtype) if bk else torch.randn(K, N, device=dev, dtype=dtype)
            if splitk:
                o = torch.zeros(M, N, device=dev, dtype=torch.float32)
                kb, tiles = (K + 63) // 64, ((M + 127) // 128) * ((N + 127) // 128)
                split = max(1, min(kb // 8, (4 * 148 + tiles - 1) // tiles))
                fn = lambda: ops.gemm([A], [Bm], M, N, K, b_kmajor=bk, out0=o, accumulate=True, split_k=split)
            else:
                o = torch.empty(M, N, device=dev, dtype=dtype)
                fn = lambda: ops.gemm([A], [Bm], M, N, K, b_kmajor=bk, out0=o)
            lib = (lambda: torch.matmul(A, Bm.t())) if bk else (lambda: torch.matmul(A, Bm))
        t = timed_isolated(fn, iters, scratch)
        tl = timed_isolated(lib, iters, scratch)
        rows.append({"name": name, "M": M, "N": N, "K": K, "us": round(t, 2), "TFLOPs": round(2 * M * N * K / t / 1e6, 1),
                     "cublas_us": round(tl, 2), "cublas_TFLOPs": round(2 * M * N * K / tl / 1e6, 1)})
        del A, Bm, o
    return rows


def adapter_subrun(dev, replays=20):
    """The adapter alone at the step's size (M = 2 * 8 * 77 = 1232 rows, shipped 6.03 M MLP), forward + backward inside a CUDA
    graph: this library's module (LayerNorm kernel + tcgen05 GEMM chain with fused GELU / GELU' / bias / pooling epilogues,
    bf16 operands, fp32 master weights) against the same math in stock PyTorch under bf16 autocast (cuBLAS GEMMs + ATen
    LayerNorm / GELU / mean kernels -- what the reference's `MLP` runs, train_sdxl_zh.py:58-67).  us per forward+backward."""
    import torch.nn as nn
    from pea_diffusion_b200 import adapter as A

    class TorchMLP(nn.Module):      # the reference's module structure, restated (plain PyTorch baseline, not the oracle)
        def __init__(self):
            super().__init__()
            self.layernorm = nn.LayerNorm(1024)
            self.projector = nn.Sequential(nn.Linear(1024, 1024, bias=False), nn.GELU(), nn.Linear(1024, 1024, bias=False), nn.GELU(),
                                           nn.Linear(1024, 1280, bias=False))
            self.fc = nn.Linear(1280, 2048)

        def forward(self, x):
            x = self.projector(self.layernorm(x))
            return torch.mean(x, 1), self.fc(nn.functional.gelu(x))

    torch.manual_seed(0)
    x = torch.randn(16, 77, 1024, device=dev)
    res = {}
    pdl_default = A.PDL_CHAIN
    for name in ("library", "library_plain_launches", "pytorch_bf16_autocast"):
        A.PDL_CHAIN = pdl_default and name == "library"     # second arm: the same chain without PEA_LAUNCH_PDL
        m = (A.MLP(1024, 1280, 1024, 2048, False, precision="bf16") if name.startswith("library") else TorchMLP()).to(dev)

        def fb():
            for p in m.parameters():
                p.grad = None
            if name.startswith("library"):
                x1, x2 = m(x)
            else:
                with torch.autocast("cuda", dtype=torch.bfloat16):
                    x1, x2 = m(x)
            (x1.float().sum() + x2.float().square().mean()).backward()

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                fb()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            fb()
        res[name + "_us"] = time_replays(g, replays) * 1e3
        del g
    A.PDL_CHAIN = pdl_default
    res["flops_fwd_bwd"] = 6 * 1232 * 6029312
    res["library_TFLOPs"] = res["flops_fwd_bwd"] / res["library_us"] / 1e6
    res["how"] = "CUDA-graph replay of one forward+backward (loss = sum(x1) + mean(x2^2)), median of %d, rows M = 1232" % replays
    return res


def infer_subrun(dev, steps=30):
    """BASELINE.json configs[4] (tests/test_sdxl_zh.py:299-406): SDXL shell fp16, UNet batch 8 (4 prompts x CFG), `steps`
    forward-only calls with a constant context; scheduler math replaced by `latents -= 0.01 * eps` (out of scope).
    Fused processors with the loop-invariant K|V cache vs the stock processors on the same weights."""
    from pea_diffusion_b200 import _lib, attn2
    from pea_diffusion_b200.shell.unet import StockAttnProcessor, UNet2DConditionShell, UNetConfig
    cfg = UNetConfig.sdxl()
    torch.manual_seed(0)
    with torch.device(dev):
        unet = UNet2DConditionShell(cfg)
    unet = unet.half().requires_grad_(False)
    g = torch.Generator(device=dev).manual_seed(1)
    ctx = torch.randn(8, 77, 2048, device=dev, dtype=torch.float16, generator=g)
    kw = {"text_embeds": torch.randn(8, 1280, device=dev, dtype=torch.float16, generator=g),
          "time_ids": torch.tensor([[1024.0, 1024, 0, 0, 1024, 1024]] * 8, device=dev)}
    lat0 = torch.randn(8, 4, 128, 128, device=dev, dtype=torch.float16, generator=g)
    tsteps = [torch.tensor([999 - 33 * i], device=dev) for i in range(steps)]
    res, outs = {}, {}
    for arm in ("stock", "fused"):
        unet.set_attn_processor(StockAttnProcessor())
        if arm == "fused":
            attn2.install(unet, cache_kv=True)

        def loop():
            lat = lat0.clone()
            with torch.no_grad():
                for i in range(steps):
                    lat = lat - 0.01 * unet(lat, tsteps[i], ctx, added_cond_kwargs=kw)[0]
            return lat

        loop()
        torch.cuda.synchronize()
        n0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); outs[arm] = loop(); e1.record()
        torch.cuda.synchronize()
        res[arm] = {"ms_per_unet_call": e0.elapsed_time(e1) / steps, "images_per_s": 4 / (e0.elapsed_time(e1) * 1e-3),
                    "library_launches": _lib.launch_count() - n0}
    ref = outs["stock"].float()
    res["rel_err_vs_stock_after_%d_steps" % steps] = float((outs["fused"].float() - ref).norm() / ref.norm())
    res["workload"] = f"SDXL shell fp16, UNet batch 8, {steps} forward-only calls, constant [8,77,2048] context, eager launches"
    del unet
    return res


# ------------------------------------------------------------------------------------------------- our arm
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
    if args.workload == "infer":
        if rank == 0:
            clocks = ClockSampler(local)
            res = infer_subrun(dev, steps=30)
            emit({"metric": "SDXL-1024 inference UNet calls/s (30-step denoise loop, UNet batch 8)", "unit": "calls/s",
                  "value": 1e3 / res["fused"]["ms_per_unet_call"], "n_gpus": 1, "steps": 30, "warmup": 30, "ms_per_step": res["fused"]["ms_per_unet_call"],
                  "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "fp16", "data": "synthetic",
                  "config": {"workload": res["workload"]}, "infer": res, "clocks": clocks.stop(),
                  "gpu_launches": res["fused"]["library_launches"]})
        return
    if world > 1:
        # stdout carries exactly ONE JSON line: NCCL's banner / debug output (stdout by default) goes to stderr
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    from pea_diffusion_b200 import _lib, adapter as A, attn2, ops
    from pea_diffusion_b200.shell.unet import UNet2DConditionShell, UNetConfig
    from pea_diffusion_b200.train import KDTrainer, parameter_checksums
    import copy

    act = torch.bfloat16 if args.dtype == "bf16" else torch.float16
    ad_prec = args.adapter_precision or args.dtype
    cfg = UNetConfig.tiny("sdxl") if args.tiny else UNetConfig.sdxl()
    in_dim, hidden = (64, 64) if args.tiny else (1024, 1024)
    latent = 16 if args.tiny else args.latent
    pooled = cfg.projection_class_embeddings_input_dim - 6 * cfg.addition_time_embed_dim
    torch.manual_seed(0)  # identical frozen weights on every rank (the adapter is broadcast from rank 0 by KDTrainer anyway)
    with torch.device(dev):
        student = UNet2DConditionShell(cfg)
    student = student.to(act)
    teacher = copy.deepcopy(student)
    attn2.install(student, freeze=True)
    attn2.install(teacher, freeze=True)
    torch.manual_seed(rank)   # deliberately DIFFERENT adapter init per rank: the trainer's broadcast has to make them equal
    adapter = A.MLP(in_dim, pooled, hidden, cfg.cross_attention_dim, False, precision=ad_prec).to(dev)
    graph_mode = not args.no_graph

    def make_trainer(overlap):
        return KDTrainer(student, teacher, adapter, "sdxl", lr=1e-5, cuda_graph=graph_mode, overlap_teacher=overlap,
                         comm_dtype=torch.bfloat16 if args.comm_dtype == "bf16" else None, capture_comm=args.capture_comm,
                         loss_scale=65536.0 if args.dtype == "fp16" else None, dynamic_loss_scale=args.dtype == "fp16")

    trainer = make_trainer(bool(args.overlap_teacher))
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

    log('models built; warm-up (graph capture)')
    warm = max(args.warmup, 3)
    for _ in range(warm):
        trainer.step(dbatch)
    torch.cuda.synchronize()

    log('timed region')
    # ---- value: inputs resident in HBM ------------------------------------------------------------------------
    clocks = ClockSampler(local) if rank == 0 else None
    ops.TIMER.reset(not graph_mode)
    n0 = _lib.launch_count()
    torch.cuda.nvtx.range_push("pea_timed")
    ms_step = timed(lambda: trainer.step(dbatch), args.steps)
    torch.cuda.nvtx.range_pop()
    ops.TIMER.enabled = False
    launches_per_step = (_lib.launch_count() - n0) / args.steps   # eager mode; a graph replay re-launches what the capture recorded
    ksum, prof_steps = ops.TIMER.summary(), args.steps
    ops.TIMER.reset(False)
    timing_note = "CUDA events around every launch inside the timed eager steps"
    ev_overhead_us = None
    log('timed region done: %.2f ms/step' % ms_step)
    if graph_mode and not args.timed_only:
        # Per-kernel times INSIDE a replayed step: the same forward+backward captured once more with event-record nodes
        # around every library launch; serial schedule (no teacher/student overlap) so each kernel has the GPU to itself.
        overlap_stream, trainer.teacher_stream = trainer.teacher_stream, None
        try:
            prof_steps = min(args.steps, 5)
            ksum, launches_per_step, _keep = profile_in_graph(trainer, prof_steps)
            del _keep
            timing_note = ("CUDA events recorded as event-record nodes around every library launch inside %d replays of a CUDA graph "
                           "of the step's forward+backward (captured after the timed region, serial schedule: no teacher/student overlap, "
                           "no collectives), so the brackets include the node-to-node launch latency" % prof_steps)
            log('in-graph per-kernel profile done')
            ev_overhead_us = event_pair_overhead_us(dev)
        finally:
            trainer.teacher_stream = overlap_stream
        torch.cuda.empty_cache()

    log('e2e loop')
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
    h2d = sum(v.numel() * v.element_size() for v in hbatch.values())

    log('strong-scaling sub-run')
    # ---- strong scaling (configs[2]): global batch 64 -> 64/W per rank in micro-batches of `--batch` ------------------
    strong = None
    if args.global_batch and not args.timed_only and args.global_batch % (world * B) == 0:
        micro = args.global_batch // (world * B)
        s_steps = max(2, min(args.steps, 3))
        trainer.step_accumulate([dbatch] * micro)
        ms_strong = timed(lambda: trainer.step_accumulate([dbatch] * micro), s_steps)
        strong = {"global_batch": args.global_batch, "per_rank_batch": micro * B, "micro_batches_per_step": micro, "steps": s_steps,
                  "ms_per_step": ms_strong, "value": args.global_batch / (ms_strong * 1e-3), "unit": UNIT, "scaling": "strong",
                  "note": "one optimizer step per global batch; gradients of the micro-batches accumulated on the device"}
    log('dp check')
    clk = clocks.stop() if clocks else None
    same, sums = parameter_checksums(adapter)
    loss_now = [float(v) for v in trainer.step(dbatch).float().cpu()]

    def comm_in_graph_subrun(line):
        """World > 1: the same step with the two bucket all-reduces CAPTURED in the step's graph (KDTrainer(capture_comm=True)):
        ms/step timed like the headline, ranks compared bitwise afterwards.  Every rank runs it; a watchdog ends the process
        (rank 0 prints the line it already has) should a collective hang."""
        import threading
        finished = threading.Event()

        def bail():
            if finished.wait(300):
                return
            if rank == 0 and line is not None:
                line["comm_in_graph"] = {"error": "timed out after 300 s"}
                emit(line)
            os._exit(0)

        threading.Thread(target=bail, daemon=True).start()
        res = None
        try:
            trainer.close()                      # drops the graph without collectives and the taps / sink
            args.capture_comm = True
            t2 = make_trainer(bool(args.overlap_teacher))
            for _ in range(3):
                t2.step(dbatch)
            ms = timed(lambda: t2.step(dbatch), args.steps)
            same2, _ = parameter_checksums(adapter)
            t2.close()                           # the graph must go before the communicator does
            res = {"ms_per_step": ms, "value": world * B / (ms * 1e-3), "unit": UNIT, "dp_params_identical": bool(same2),
                   "how": "forward + backward + both bucket all-reduces in ONE CUDA graph (forked branch launched from the adapter's "
                          "backward), then the fused AdamW launch; %d steps, max over ranks" % args.steps}
        except Exception as exc:  # noqa: BLE001
            res = {"error": str(exc).splitlines()[0][:200]}
        finally:
            args.capture_comm = False
            finished.set()
        return res

    want_cig = world > 1 and graph_mode and not args.capture_comm and not args.timed_only and not args.tiny
    if world > 1:
        dist.barrier()
    if rank != 0:
        if world > 1:
            if want_cig:
                comm_in_graph_subrun(None)
            safe_destroy_process_group()
        return

    pk, pk_kind = peaks()
    at = {k: ksum.get(k, {"launches": 0, "ms": 0.0, "bytes": 0, "flops": 0}) for k in ("attn2_fwd", "attn2_bwd", "kd_mse", "gemm")}
    a_ms = at["attn2_fwd"]["ms"] + at["attn2_bwd"]["ms"]
    a_bytes = at["attn2_fwd"]["bytes"] + at["attn2_bwd"]["bytes"]
    a_launch = at["attn2_fwd"]["launches"] + at["attn2_bwd"]["launches"]
    achieved = a_bytes / (a_ms * 1e-3) / 1e9 if a_ms > 0 else 0.0
    # DRAM traffic per launch (read + write) of the same launch mix, from the committed ncu --set full capture of the two
    # shapes a step launches; null on other workloads
    traffic, traffic_src = None, None
    for cand in ("r2_attn2_traffic.json", "r1_attn2_traffic.json"):
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", cand)))
            if not args.tiny and args.latent == 128 and args.batch == 8:
                traffic = tj["traffic_bytes_per_launch_step_mix"]
                traffic_src = "profiles/%s (ncu --set full, bytes per launch, step mix; algorithmic: %d)" % (cand, tj["algorithmic_bytes_per_launch_step_mix"])
            break
        except Exception:
            continue
    per_kernel = {k: {"launches_per_step": v["launches"] / prof_steps, "ms_per_step": v["ms"] / prof_steps,
                      "us_per_launch": (v["ms"] * 1e3 / v["launches"] if v["launches"] else None),
                      "GBps": (v["bytes"] / (v["ms"] * 1e-3) / 1e9 if v["ms"] > 0 else 0.0),
                      "TFLOPs": (v["flops"] / (v["ms"] * 1e-3) / 1e12 if v["ms"] > 0 else 0.0)} for k, v in at.items()}
    roof = {"bound": "hbm", "kernel": "attn2 fused forward + backward (all 70 layers x student fwd/bwd + teacher fwd)",
            "achieved": achieved, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": achieved / pk["hbm_gbs"],
            "frac_of_nominal_8TBps": achieved / 8000.0,
            "traffic": traffic, "traffic_source": traffic_src, "peak_source": pk_kind,
            "algorithmic_bytes_per_step": a_bytes / prof_steps if prof_steps else None,
            "bytes_formula": "SURVEY.md 8d: fwd 2BNCe + 2BLCe; bwd 3BNCe + 4BLCe (e = 2)",
            "event_pair_overhead_us": ev_overhead_us, "timing": timing_note, "per_kernel": per_kernel}
    value = world * B / (ms_step * 1e-3)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": warm,
        "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.dtype,
        "data": "synthetic", "config": workload_config(args, world),
        "execution": {"adapter_precision": ad_prec, "comm_dtype": args.comm_dtype,
                      "step": ("ONE CUDA graph per step: forward + backward" + (" + bucketed NCCL all-reduce (captured)" if world > 1 and args.capture_comm else "") +
                               (", then two NCCL bucket all-reduces of the flat gradient buffer" if world > 1 and not args.capture_comm else "") +
                               ", then the fused AdamW launch" if graph_mode else "eager") +
                              ("; teacher forward on a parallel graph branch" if args.overlap_teacher else "")},
        "e2e": {"value": world * B / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 16,
                "ms_per_step": ms_e2e},
        "gpu_launches": int(round(launches_per_step * args.steps)),
        "clocks": clk,
        "roofline": roof,
        "strong": strong,
        "dp_params_identical": bool(same), "dp_param_checksums": sums, "loss": loss_now,
    }
    if want_cig:
        log('collectives-in-graph sub-run')
        line["comm_in_graph"] = comm_in_graph_subrun(line)
    if world == 1 and not args.timed_only:
        # free the training state before the isolated-kernel sweeps
        trainer.close()
        del trainer, student, teacher
        torch.cuda.empty_cache()
        if not args.no_sweep and not args.tiny:
            clocks2 = ClockSampler(local)
            try:
                log('attn2 / gemm sweeps')
                roof["in_chain_marginal"] = chain_marginal(dev, pk["hbm_gbs"])
                roof["sweep"] = attn2_sweep(dev, pk["hbm_gbs"])
                roof["sweep_gemm"] = gemm_sweep(dev)
                roof["adapter_fwd_bwd"] = adapter_subrun(dev)
                roof["sweep_how"] = ("each launch alone, 5 warm-up + 20 timed iterations, CUDA events on the launching stream, L2 flushed by "
                                     "reading 256 MB between iterations; median; fractions of the measured HBM peak")
            except Exception as exc:
                roof["sweep_error"] = str(exc).splitlines()[0][:200]
            roof["sweep_clocks"] = clocks2.stop()
        if not args.no_infer and not args.tiny:
            try:
                log('inference sub-run')
                line["infer"] = infer_subrun(dev)
            except Exception as exc:
                line["infer"] = {"error": str(exc).splitlines()[0][:200]}
        if not args.no_cpu_baseline:
            try:
                log('cpu baseline sample')
                cl, cb = (16, 2) if args.tiny else (32, 1)
                sec, cores = cpu_reference_sample(2, 1, cl, cb, args.tiny)
                scale = (latent / cl) ** 2
                line["cpu_baseline"] = {"value": cb / sec / scale, "unit": UNIT, "cores": cores, "kind": "port",
                                        "sample": f"oracle KD step fp32, SDXL shell, batch {cb}, {cl}x{cl} latent, {sec:.2f} s/step; samples/s "
                                                  f"divided by {scale:.0f} (token ratio to {latent}x{latent}; extrapolation)"}
            except Exception as e:  # the GPU numbers stand on their own
                line["cpu_baseline"] = {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": f"failed: {e}"}
    emit(line)
    if world > 1:
        safe_destroy_process_group()


if __name__ == "__main__":
    main()
