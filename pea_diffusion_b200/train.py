"""The KD training step of PEA-Diffusion on the B200 kernels, and its data-parallel wrapper.

Mirrors ``StableDiffusion.training_step`` (train_sdxl_zh.py:305-449 / train_sd_zh.py:150-272) from the point where the
reference has latents, noise, timesteps and text-encoder states in hand (VAE / text encoders / data loading are out
of the hot-path scope):

    4  adapter(cond), adapter(uncond)                     :383-384   -> one batched call on 2B rows
    6  CFG-dropout select with a shared mask              :392-395
    7  student UNet (attn2 -> PEAAttn2Processor)          :397
    9  teacher UNet under no_grad                         :410-415
    8/10/11  base + output-level + feature-level KD loss  :400-441   -> one fused loss+gradient pass, no host sync
    DP: mean all-reduce of the adapter gradients over NCCL through ONE flat buffer the dW GEMMs write into (two buckets,
        the head's first; the reference gets this from DeepSpeed ZeRO-1, train_sdxl_zh.sh:22-32), then an identical AdamW
        step per rank.
"""
from __future__ import annotations

from typing import Dict, List, Mapping, Optional, Tuple

import torch
import torch.distributed as dist
import torch.nn as nn

from . import ops
from .kd_loss import kd_loss

Tensor = torch.Tensor


class _CfgSelectFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, mask_u8: Tensor, uncond: Tensor, cond: Tensor):
        ctx.save_for_backward(mask_u8)
        return ops.cfg_select(mask_u8, uncond, cond)

    @staticmethod
    def backward(ctx, g: Tensor):
        (mask_u8,) = ctx.saved_tensors
        z = torch.zeros_like(g)
        g = g.contiguous()
        # uncond rows get g where mask else 0; cond rows the complement
        return None, ops.cfg_select(mask_u8, g, z), ops.cfg_select(mask_u8, z, g)


def cfg_dropout_select(prompt_mask: Tensor, uncond: Tensor, cond: Tensor) -> Tensor:
    """``torch.where(prompt_mask[:, None, None], uncond, cond)`` (train_sdxl_zh.py:392-395) with gradient routing."""
    return _CfgSelectFn.apply(prompt_mask.to(torch.uint8).contiguous(), uncond, cond)


class KDTaps:
    """Forward hooks of train_sdxl_zh.py:69-84,153-156: down blocks store ``output[0]``, mid/up blocks ``output``
    (references, no copies).  Order d0..dn-1, m, u0..un-1 is the summation order of :424-438."""

    def __init__(self, unet: nn.Module):
        self.store: Dict[str, Tensor] = {}
        self.handles = []
        n = len(unet.down_blocks)
        for i, blk in enumerate(unet.down_blocks):
            self.handles.append(blk.register_forward_hook(lambda m, a, out, k=f"d{i}": self.store.__setitem__(k, out[0])))
        self.handles.append(unet.mid_block.register_forward_hook(lambda m, a, out: self.store.__setitem__("m", out)))
        for i, blk in enumerate(unet.up_blocks):
            self.handles.append(blk.register_forward_hook(lambda m, a, out, k=f"u{i}": self.store.__setitem__(k, out)))
        self.order = [f"d{i}" for i in range(n)] + ["m"] + [f"u{i}" for i in range(len(unet.up_blocks))]

    def tensors(self) -> List[Tensor]:
        return [self.store[k] for k in self.order]

    def clear(self) -> None:
        self.store.clear()

    def remove(self) -> None:
        for h in self.handles:
            h.remove()
        self.handles = []


def kd_forward(student: nn.Module, teacher: nn.Module, adapter: nn.Module, batch: Mapping[str, Tensor], taps_s: KDTaps,
               taps_t: KDTaps, variant: str = "sdxl", feature_weight: float = 0.1,
               skip_nonfinite: Optional[bool] = None, teacher_stream: Optional["torch.cuda.Stream"] = None,
               grad_prescale: float = 1.0) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
    """Forward graph of one KD step on device tensors; returns (loss, L0, L1, L2) without any host sync.

    ``batch`` keys: noisy_latents, noise, timesteps, enc, enc_uncond, teacher_ctx,
    teacher_ctx_uncond, zh_or_not, prompt_mask, and for SDXL teacher_pooled, time_ids.

    ``teacher_stream``: run the teacher's forward (no_grad, independent of the adapter: train_sdxl_zh.py:410-415) on that
    stream, concurrently with the student's forward, and join before the loss.  Inside a CUDA-graph capture this becomes
    a parallel branch of the graph.  The result is identical; only the schedule changes.
    """
    if skip_nonfinite is None:
        skip_nonfinite = variant == "sd"
    act_dtype = batch["noisy_latents"].dtype
    enc, enc_u = batch["enc"], batch["enc_uncond"]
    B = enc.shape[0]
    both = torch.cat([enc, enc_u], dim=0)  # one 2B-row adapter pass instead of two (mean-pool is per sample: exact)
    mask = batch["prompt_mask"]
    if variant == "sdxl":
        x1, x2 = adapter(both)
        ctx = cfg_dropout_select(mask, x2[B:].to(act_dtype), x2[:B].to(act_dtype))
        kw_s = {"text_embeds": x1[:B].to(act_dtype), "time_ids": batch["time_ids"]}  # pooled stays conditional (:412-413)
        kw_t = {"text_embeds": batch["teacher_pooled"], "time_ids": batch["time_ids"]}
    else:
        x2 = adapter(both)
        ctx = cfg_dropout_select(mask, x2[B:].to(act_dtype), x2[:B].to(act_dtype))
        kw_s = kw_t = None
    taps_s.clear()
    taps_t.clear()

    def run_teacher():
        with torch.no_grad():
            t_ctx = cfg_dropout_select(mask, batch["teacher_ctx_uncond"], batch["teacher_ctx"])
            return teacher(batch["noisy_latents"], batch["timesteps"], t_ctx, added_cond_kwargs=kw_t)[0]

    if teacher_stream is not None:
        main = torch.cuda.current_stream()
        teacher_stream.wait_stream(main)           # batch tensors (and the previous step's readers of recycled memory)
        with torch.cuda.stream(teacher_stream):
            pred_t = run_teacher()
        pred = student(batch["noisy_latents"], batch["timesteps"], ctx, added_cond_kwargs=kw_s)[0]
        main.wait_stream(teacher_stream)           # join: the loss reads the teacher's prediction and taps
        # No record_stream needed: teacher_stream allocates nothing more in this step, and its next use (next step) starts
        # with wait_stream(main), i.e. after every main-stream reader of the tensors it produced here.
    else:
        pred = student(batch["noisy_latents"], batch["timesteps"], ctx, added_cond_kwargs=kw_s)[0]
        pred_t = run_teacher()
    return kd_loss(pred, batch["noise"], pred_t, taps_s.tensors(), taps_t.tensors(), batch["zh_or_not"],
                   feature_weight=feature_weight, skip_nonfinite=skip_nonfinite, grad_prescale=grad_prescale)


class GradAllReducer:
    """Mean all-reduce of a parameter list's gradients through ONE flat buffer (what DeepSpeed ZeRO-1 / DDP give the
    reference, train_sdxl_zh.sh:22-32,87; 6 033 408 elements for the shipped adapter).

    ``param.grad`` of every parameter is a view of ``flat`` (fp32).  Two ways gradients get there:
      * sink protocol (this library's adapter, ``attach(adapter)``): the adapter's backward writes dW / db straight into
        ``slot(param)`` -- the dW GEMM's epilogue stores into the communication buffer, in ``comm_dtype`` -- and calls
        ``mark_ready(param)``; a bucket's collective is launched on the side stream the moment its last gradient is
        written (bucket 0 = the head ``fc.*``, produced first, so its 10.5 MB are in flight while the rest of the adapter
        backward runs).  Works inside a CUDA-graph capture: the collectives become a forked branch of the graph.
      * generic modules: autograd's post-accumulate hooks copy each gradient into its slot; the last one launches a single
        collective over the whole buffer.
    ``finish()`` joins the side stream (and, for a 16-bit ``comm_dtype``, widens the reduced values into ``flat``) before
    the optimizer step.  ``comm_dtype=torch.bfloat16`` halves the payload (12 MB instead of 24 MB).
    Backend-agnostic: NCCL on GPUs (``ReduceOp.AVG``), gloo in the CPU tests (SUM then scale).
    """

    def __init__(self, params, process_group=None, hooks: bool = True, comm_dtype: Optional[torch.dtype] = None,
                 order: Optional[List[Tensor]] = None, bucket_split: Optional[int] = None):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self._pending: List = []
        self._hooks = []
        on_gpu = len(self.params) > 0 and all(p.is_cuda for p in self.params)
        self._comm_stream = torch.cuda.Stream() if (self.world > 1 and on_gpu) else None
        layout = list(order) if order is not None else list(self.params)
        assert {id(p) for p in layout} == {id(p) for p in self.params}, "order must be a permutation of the trainable parameters"
        dev = layout[0].device if layout else torch.device("cpu")
        total = sum(p.numel() for p in layout)
        self.flat = torch.zeros(total, device=dev, dtype=torch.float32)
        self.comm_dtype = comm_dtype or torch.float32
        self.comm = self.flat if self.comm_dtype == torch.float32 else torch.zeros(total, device=dev, dtype=self.comm_dtype)
        self._span: Dict[int, Tuple[int, int]] = {}
        off = 0
        for p in layout:
            self._span[id(p)] = (off, off + p.numel())
            off += p.numel()
        # buckets: [0, split) and [split, total); split = end of the first `bucket_split` parameters of the layout
        split = sum(p.numel() for p in layout[:bucket_split]) if bucket_split else 0
        self._buckets = [(0, split), (split, total)] if 0 < split < total else [(0, total)]
        self._bucket_of = {id(p): (0 if self._span[id(p)][0] < self._buckets[0][1] else len(self._buckets) - 1) for p in layout}
        self._need = [sum(1 for p in layout if self._bucket_of[id(p)] == b) for b in range(len(self._buckets))]
        self._got = [0] * len(self._buckets)
        self._layout = layout
        self._views_bound = False
        self.bytes_per_step = total * self.comm.element_size()
        use_avg = self.world > 1 and on_gpu and dist.get_backend(process_group) == "nccl"
        self._op = dist.ReduceOp.AVG if use_avg else dist.ReduceOp.SUM
        self._post_scale = 1.0 if (use_avg or self.world == 1) else 1.0 / self.world
        if hooks:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._hook))

    # -- sink protocol ----------------------------------------------------------------------------------------------
    def attach(self, module: nn.Module) -> None:
        """Make ``module`` (an adapter of this library) write its parameter gradients straight into the flat buffer."""
        self.bind_views()
        module._grad_sink = self
        for h in self._hooks:        # the sink replaces the hooks for these parameters
            h.remove()
        self._hooks = []

    def bind_views(self) -> None:
        """``param.grad`` := fp32 view of the flat buffer (kept across steps: zero_grad(set_to_none=True) would detach them)."""
        for p in self._layout:
            a, b = self._span[id(p)]
            p.grad = self.flat[a:b].view_as(p)
        self._views_bound = True

    def slot(self, p: Tensor) -> Optional[Tensor]:
        span = self._span.get(id(p))
        return None if span is None else self.comm[span[0]:span[1]].view_as(p)

    def mark_ready(self, p: Tensor) -> None:
        b = self._bucket_of[id(p)]
        self._got[b] += 1
        if self._got[b] == self._need[b]:
            self._got[b] = 0
            self._launch(b)

    # -- generic path: autograd hooks ----------------------------------------------------------------------------------
    def _hook(self, p: Tensor) -> None:
        a, b = self._span[id(p)]
        view = self.flat[a:b].view_as(p)
        if p.grad is not None and p.grad.data_ptr() != view.data_ptr():
            self.comm[a:b].view_as(p).copy_(p.grad)
            p.grad = view
        elif self.comm is not self.flat:
            self.comm[a:b].view_as(p).copy_(p.grad)
        self.mark_ready(p)

    def reduce_now(self) -> None:
        """Explicit variant (CUDA-graph replays of a graph captured WITHOUT the collectives): reduce every bucket."""
        for b in range(len(self._buckets)):
            self._got[b] = 0
            self._launch(b)

    def _launch(self, b: int) -> None:
        if self.world == 1:
            return
        lo, hi = self._buckets[b]
        buf = self.comm[lo:hi]
        if self._comm_stream is not None:
            self._comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._comm_stream):
                self._pending.append(dist.all_reduce(buf, op=self._op, group=self.group, async_op=True))
        else:
            self._pending.append(dist.all_reduce(buf, op=self._op, group=self.group, async_op=True))

    def finish(self) -> None:
        if self._pending:
            for w in self._pending:
                w.wait()
            self._pending.clear()
            if self._comm_stream is not None:     # only when something was launched there: inside a CUDA-graph capture that
                torch.cuda.current_stream().wait_stream(self._comm_stream)   # carries no collectives the side stream is not captured
        if self.comm is not self.flat:
            self.flat.copy_(self.comm)          # widen the reduced 16-bit payload into the fp32 gradients
        if self._post_scale != 1.0:
            self.flat.mul_(self._post_scale)

    def close(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []


def broadcast_parameters(module: nn.Module, process_group=None, src: int = 0) -> None:
    """Every rank starts from rank ``src``'s adapter (what DeepSpeed / DDP do at wrap time for the reference): without it
    a per-rank default init, or a checkpoint loaded on one rank only, silently trains W different adapters."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(process_group) == 1:
        return
    gsrc = dist.get_global_rank(process_group, src) if process_group is not None else src
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()):
            dist.broadcast(t.data, src=gsrc, group=process_group)


def parameter_checksums(module: nn.Module, process_group=None) -> Tuple[bool, List[float]]:
    """-> (identical on every rank?, per-rank fp64 checksums).  Bitwise: compares the gathered raw parameter bytes'
    integer sums, not a tolerance."""
    flat = torch.cat([p.detach().reshape(-1).float() for p in module.parameters()])
    ints = flat.view(torch.int32).to(torch.int64)
    sig = torch.stack([ints.sum(), (ints * (torch.arange(ints.numel(), device=ints.device) % 8191 + 1)).sum()])
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(process_group) == 1:
        return True, [float(flat.double().sum())]
    world = dist.get_world_size(process_group)
    sigs = [torch.empty_like(sig) for _ in range(world)]
    dist.all_gather(sigs, sig, group=process_group)
    sums = torch.zeros(world, device=flat.device, dtype=torch.float64)
    sums[dist.get_rank(process_group)] = flat.double().sum()
    dist.all_reduce(sums, group=process_group)
    return all(torch.equal(sigs[0], s_) for s_ in sigs), [float(v) for v in sums.cpu()]


class KDTrainer:
    """Owns the adapter's optimizer and (when ``torch.distributed`` is initialised) the gradient all-reduce.

    One process per GPU; every rank holds full replicas of both UNets and the adapter (broadcast from rank 0 at
    construction); the only exchange per step is the mean all-reduce of the adapter gradients (6.03 M parameters, one flat
    buffer, ``GradAllReducer``), then an identical AdamW step on every rank (replaces the reference's ZeRO-1 sharding,
    train_sdxl_zh.sh:22-31).

    ``cuda_graph=True`` captures forward + backward of the step (adapter, both UNets, loss: every kernel of this library
    and of the PyTorch shell) into ONE CUDA graph on the first call and replays it afterwards; batches are copied into the
    graph's static input buffers.  After the replay two NCCL launches reduce the flat gradient buffer (bucket by bucket, on
    the side stream) and the fused AdamW step follows; shapes must stay fixed.  ``capture_comm=True`` puts the collectives
    into the graph as a forked branch (launched from inside the adapter's backward, joined before the capture ends): the
    replay then delivers rank-averaged gradients.  Call ``close()`` before destroying the process group in that mode.

    ``loss_scale`` (fp16 training, the reference's ``--precision 16``, train_sdxl_zh.sh:89): the loss is scaled before
    backward -- the fused KD pass applies the scale before rounding its gradients to fp16 --, gradients are unscaled and
    checked on the device, a non-finite step is skipped by the fused AdamW (``found_inf``), and with
    ``dynamic_loss_scale`` the scale is halved / doubled on the device (no host sync anywhere).
    """

    def __init__(self, student: nn.Module, teacher: nn.Module, adapter: nn.Module, variant: str = "sdxl", lr: float = 1e-5,
                 betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, feature_weight: float = 0.1,
                 process_group=None, cuda_graph: bool = False, overlap_teacher: bool = False,
                 comm_dtype: Optional[torch.dtype] = None, capture_comm: bool = False,
                 loss_scale: Optional[float] = None, dynamic_loss_scale: bool = False, growth_interval: int = 2000):
        self.student, self.teacher, self.adapter, self.variant = student, teacher, adapter, variant
        self.feature_weight = feature_weight
        self.taps_s, self.taps_t = KDTaps(student), KDTaps(teacher)
        broadcast_parameters(adapter, process_group)
        params = [p for p in adapter.parameters() if p.requires_grad]
        fused = all(p.is_cuda for p in params)
        # utils/model_utils.py:52-71: AdamW on the adapter only; the --weight_decay flag never reaches the param
        # group, so the effective decay is 0
        self.opt = torch.optim.AdamW(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, fused=fused)
        self.cuda_graph = cuda_graph
        # teacher forward on its own stream, concurrent with the student forward (a parallel branch under cuda_graph)
        self.teacher_stream = torch.cuda.Stream() if overlap_teacher and fused else None
        # flat layout in the order the adapter's backward produces the gradients: head (fc.*) first, LayerNorm last
        named = [(n, p) for n, p in adapter.named_parameters() if p.requires_grad]
        head = [p for n, p in named if n.startswith(("fc.", "fc3."))]
        rest = [p for n, p in reversed(named) if not n.startswith(("fc.", "fc3."))]
        self.reducer = GradAllReducer(params, process_group, hooks=False, comm_dtype=comm_dtype, order=head + rest,
                                      bucket_split=len(head) if head else None)
        self.reducer.attach(adapter)
        self.capture_comm = capture_comm
        self._graph = None
        self._static_in: Dict[str, Tensor] = {}
        self._static_out: Optional[Tensor] = None
        self.loss_scale = None if loss_scale is None else float(loss_scale)
        self.dynamic_loss_scale = dynamic_loss_scale and loss_scale is not None
        if self.loss_scale is not None:
            dev = params[0].device
            # 0-dim tensors: what torch's fused AdamW (found_inf) and the amp helpers expect
            self._scale = torch.full((), self.loss_scale, device=dev, dtype=torch.float32)
            self._growth = torch.zeros((), device=dev, dtype=torch.int32)
            self._found_inf = torch.zeros((), device=dev, dtype=torch.float32)
            self._growth_interval = growth_interval

    def _forward_backward(self, batch: Mapping[str, Tensor], reduce: bool = True) -> Tensor:
        loss, l0, l1, l2 = kd_forward(self.student, self.teacher, self.adapter, batch, self.taps_s, self.taps_t, self.variant,
                                      self.feature_weight, teacher_stream=self.teacher_stream,
                                      grad_prescale=self.loss_scale or 1.0)
        if self.loss_scale is None:
            loss.backward()
        else:
            loss.backward(gradient=self._scale.to(loss.dtype))
        self.taps_s.clear()
        self.taps_t.clear()
        if reduce:
            self.reducer.finish()       # joins the collectives the adapter's backward launched (inside a capture: a graph join)
        return torch.stack([loss.detach(), l0, l1, l2])

    def _capture(self, batch: Mapping[str, Tensor]) -> None:
        self._static_in = {k: v.clone() for k, v in batch.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        world = self.reducer.world
        if world > 1 and not self.capture_comm:
            self.reducer.world = 1      # the collectives stay outside the graph: issued by reduce_now() after each replay
        with torch.cuda.stream(side):   # warm-up off the capture stream: lazy initialisation, allocator, autotuners, NCCL
            for _ in range(2):
                self._forward_backward(self._static_in)
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph, stream=side, capture_error_mode="thread_local" if (world > 1 and self.capture_comm) else "global"):
            self._static_out = self._forward_backward(self._static_in)
        self.reducer.world = world

    def _optimizer_step(self) -> None:
        if self.loss_scale is not None:
            # unscale + finite check on the device; the fused AdamW skips the update when found_inf is set
            self._found_inf.zero_()
            inv = self._scale.reciprocal()
            torch._amp_foreach_non_finite_check_and_unscale_([self.reducer.flat], self._found_inf, inv)
            self.opt.found_inf = self._found_inf
            self.opt.grad_scale = None
        self.opt.step()
        if self.dynamic_loss_scale:
            torch._amp_update_scale_(self._scale, self._growth, self._found_inf, 2.0, 0.5, self._growth_interval)

    def step(self, batch: Mapping[str, Tensor]) -> Tensor:
        """forward + backward + all-reduce + AdamW.  Returns the device tensor [L, L0, L1, L2]; no host sync."""
        if self.cuda_graph:
            if self._graph is None:
                self._capture(batch)
            for k, v in batch.items():
                self._static_in[k].copy_(v, non_blocking=True)
            self._graph.replay()     # gradients (already averaged over ranks when the collectives were captured) in reducer.flat
            if self.reducer.world > 1 and not self.capture_comm:
                self.reducer.reduce_now()
                self.reducer.finish()
            self._optimizer_step()
            return self._static_out
        out = self._forward_backward(batch)
        self._optimizer_step()
        return out

    def step_accumulate(self, batches) -> Tensor:
        """One optimizer step over several micro-batches (DeepSpeed's gradient accumulation, train_sdxl_zh.sh:21-32: the
        global batch is micro-batch x accumulation x ranks): forward+backward per micro-batch, gradients summed on the
        device, averaged, then ONE AdamW step.  With the collectives captured in the graph every micro-batch is averaged
        over the ranks as it completes (mean of means = global mean for equal micro-batches)."""
        batches = list(batches)
        k = len(batches)
        if k == 1:
            return self.step(batches[0])
        flat = self.reducer.flat
        acc = None
        out = None
        for b in batches:
            if self.cuda_graph:
                if self._graph is None:
                    self._capture(b)
                for key, v in b.items():
                    self._static_in[key].copy_(v, non_blocking=True)
                self._graph.replay()
                if self.reducer.world > 1 and not self.capture_comm:
                    self.reducer.reduce_now()
                    self.reducer.finish()
                o = self._static_out
            else:
                o = self._forward_backward(b)
            acc = flat.clone() if acc is None else acc.add_(flat)
            out = o.clone() if out is None else out.add_(o)
        torch.mul(acc, 1.0 / k, out=flat)
        self._optimizer_step()
        return out.mul_(1.0 / k)

    def close(self) -> None:
        """Detach from the modules and DROP the step graph.  Required before ``dist.destroy_process_group()`` when the
        collectives were captured (``capture_comm=True``): NCCL does not tear a communicator down while a CUDA graph that
        holds its collectives is alive -- the teardown blocks (measured: tools/nccl_capture_probe.py)."""
        self.reducer.close()
        if getattr(self.adapter, "_grad_sink", None) is self.reducer:
            self.adapter._grad_sink = None
        self.taps_s.remove()
        self.taps_t.remove()
        if self._graph is not None:
            torch.cuda.synchronize()
            self._graph = None
            self._static_out = None
            self._static_in = {}
