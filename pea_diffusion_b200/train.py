"""The KD training step of PEA-Diffusion on the B200 kernels, and its data-parallel wrapper.

Mirrors ``StableDiffusion.training_step`` (train_sdxl_zh.py:305-449 / train_sd_zh.py:150-272) from the point where the
reference has latents, noise, timesteps and text-encoder states in hand (VAE / text encoders / data loading are out
of the hot-path scope):

    4  adapter(cond), adapter(uncond)                     :383-384   -> one batched call on 2B rows
    6  CFG-dropout select with a shared mask              :392-395
    7  student UNet (attn2 -> PEAAttn2Processor)          :397
    9  teacher UNet under no_grad                         :410-415
    8/10/11  base + output-level + feature-level KD loss  :400-441   -> one fused loss+gradient pass, no host sync
    DP: mean all-reduce of the 7 adapter gradients over NCCL, issued per tensor from the gradient-ready hooks (the
        reference gets this from DeepSpeed ZeRO-1, train_sdxl_zh.sh:22-32), then an identical AdamW step per rank.
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
               skip_nonfinite: Optional[bool] = None, teacher_stream: Optional["torch.cuda.Stream"] = None
               ) -> Tuple[Tensor, Tensor, Tensor, Tensor]:
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
                   feature_weight=feature_weight, skip_nonfinite=skip_nonfinite)


class GradAllReducer:
    """Mean all-reduce of a parameter list's gradients, one collective per tensor, launched from autograd's
    post-accumulate hooks (so ``fc.weight``'s reduction is in flight while the rest of the adapter backward runs) on a
    side stream; ``finish()`` joins before the optimizer step.  Backend-agnostic: NCCL on GPUs, gloo in the CPU tests.
    """

    def __init__(self, params, process_group=None, hooks: bool = True):
        self.params = [p for p in params if p.requires_grad]
        self.group = process_group
        self.world = dist.get_world_size(process_group) if dist.is_available() and dist.is_initialized() else 1
        self._pending: List = []
        self._hooks = []
        on_gpu = len(self.params) > 0 and all(p.is_cuda for p in self.params)
        self._comm_stream = torch.cuda.Stream() if (self.world > 1 and on_gpu) else None
        self.bytes_per_step = sum(p.numel() * p.element_size() for p in self.params)
        if self.world > 1 and hooks:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._grad_ready))

    def reduce_now(self) -> None:
        """Explicit variant for CUDA-graph replays (autograd hooks do not fire on replay): launch every tensor's
        all-reduce, last-produced gradient first."""
        if self.world > 1:
            for p in reversed(self.params):
                if p.grad is not None:
                    self._grad_ready(p)

    def _grad_ready(self, p: Tensor) -> None:
        if self._comm_stream is not None:
            self._comm_stream.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(self._comm_stream):
                p.grad.div_(self.world)
                self._pending.append(dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
        else:
            p.grad.div_(self.world)
            self._pending.append(dist.all_reduce(p.grad, op=dist.ReduceOp.SUM, group=self.group, async_op=True))

    def finish(self) -> None:
        for w in self._pending:
            w.wait()
        self._pending.clear()
        if self._comm_stream is not None:
            torch.cuda.current_stream().wait_stream(self._comm_stream)

    def close(self) -> None:
        for h in self._hooks:
            h.remove()
        self._hooks = []


class KDTrainer:
    """Owns the adapter's optimizer and (when ``torch.distributed`` is initialised) the gradient all-reduce.

    One process per GPU; every rank holds full replicas of both UNets and the adapter; the only exchange per step is
    the mean all-reduce of the adapter gradients (6.03 M parameters, ``GradAllReducer``), then an identical AdamW step
    on every rank (replaces the reference's ZeRO-1 sharding, train_sdxl_zh.sh:22-31).

    ``cuda_graph=True`` captures forward + backward of the step (adapter, both UNets, loss: every kernel of this
    library and of the PyTorch shell) into ONE CUDA graph on the first call and replays it afterwards; batches are
    copied into the graph's static input buffers.  The all-reduce and the fused AdamW step stay outside the graph.
    This removes the Python / launch overhead of the ~10^4 launches per step; shapes must then stay fixed.
    """

    def __init__(self, student: nn.Module, teacher: nn.Module, adapter: nn.Module, variant: str = "sdxl", lr: float = 1e-5,
                 betas=(0.9, 0.999), eps: float = 1e-8, weight_decay: float = 0.0, feature_weight: float = 0.1,
                 process_group=None, cuda_graph: bool = False, overlap_teacher: bool = False):
        self.student, self.teacher, self.adapter, self.variant = student, teacher, adapter, variant
        self.feature_weight = feature_weight
        self.taps_s, self.taps_t = KDTaps(student), KDTaps(teacher)
        params = [p for p in adapter.parameters() if p.requires_grad]
        fused = all(p.is_cuda for p in params)
        # utils/model_utils.py:52-71: AdamW on the adapter only; the --weight_decay flag never reaches the param
        # group, so the effective decay is 0
        self.opt = torch.optim.AdamW(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, fused=fused)
        self.cuda_graph = cuda_graph
        # teacher forward on its own stream, concurrent with the student forward (a parallel branch under cuda_graph)
        self.teacher_stream = torch.cuda.Stream() if overlap_teacher and fused else None
        self.reducer = GradAllReducer(params, process_group, hooks=not cuda_graph)
        self._graph = None
        self._static_in: Dict[str, Tensor] = {}
        self._static_out: Optional[Tensor] = None

    def _forward_backward(self, batch: Mapping[str, Tensor]) -> Tensor:
        loss, l0, l1, l2 = kd_forward(self.student, self.teacher, self.adapter, batch, self.taps_s, self.taps_t, self.variant,
                                      self.feature_weight, teacher_stream=self.teacher_stream)
        loss.backward()
        self.taps_s.clear()
        self.taps_t.clear()
        return torch.stack([loss.detach(), l0, l1, l2])

    def _capture(self, batch: Mapping[str, Tensor]) -> None:
        self._static_in = {k: v.clone() for k, v in batch.items()}
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):   # warm-up off the capture stream: lazy initialisation, allocator, autotuners
            for _ in range(2):
                self.opt.zero_grad(set_to_none=True)
                self._forward_backward(self._static_in)
        torch.cuda.current_stream().wait_stream(side)
        self.opt.zero_grad(set_to_none=True)
        self._graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self._graph, stream=side):   # same stream as the warm-up: AccumulateGrad nodes stay put
            self._static_out = self._forward_backward(self._static_in)

    def step(self, batch: Mapping[str, Tensor]) -> Tensor:
        """forward + backward + (all-reduce) + AdamW.  Returns the device tensor [L, L0, L1, L2]; no host sync."""
        if self.cuda_graph:
            if self._graph is None:
                self._capture(batch)
            for k, v in batch.items():
                self._static_in[k].copy_(v, non_blocking=True)
            self._graph.replay()     # gradients land in the static .grad tensors the capture created
            self.reducer.reduce_now()
            self.reducer.finish()
            self.opt.step()
            return self._static_out
        self.opt.zero_grad(set_to_none=True)
        out = self._forward_backward(batch)
        self.reducer.finish()
        self.opt.step()
        return out

    def close(self) -> None:
        self.reducer.close()
        self.taps_s.remove()
        self.taps_t.remove()
