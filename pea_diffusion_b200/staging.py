"""Host -> HBM staging of the step's batch, double-buffered on a copy stream.

The reference hands `training_step` a collated dict of CPU tensors (`utils/custom_dataset_sdxl.py:384-409`, read at
`train_sdxl_zh.py:311-397`) and lets Lightning move it to the GPU synchronously at the top of every step.  Here the
copy of batch k+1 runs on its own stream while step k computes:

    stager = BatchStager(device)
    for dev_batch in stager.prefetch(host_batches):      # host_batches: iterable of {name: CPU tensor}
        trainer.step(dev_batch)

Each slot owns pinned host buffers and device buffers of fixed addresses (allocated on first use per key / shape /
dtype), so nothing is allocated in steady state and the device side of a slot can feed a captured CUDA graph's static
inputs with plain device-to-device copies.  Ordering is by CUDA events only; the one host wait is for the slot's OWN
previous upload before its pinned buffer is overwritten (two steps old with the default depth, i.e. long finished).
"""
from __future__ import annotations

from typing import Dict, Iterable, Iterator, List, Mapping, Optional

import torch
from torch import Tensor

from ._lib import PeaError


class _Slot:
    def __init__(self) -> None:
        self.pinned: Dict[str, Tensor] = {}
        self.dev: Dict[str, Tensor] = {}
        self.copied: Optional[torch.cuda.Event] = None      # upload of this slot finished (recorded on the copy stream)
        self.consumed: Optional[torch.cuda.Event] = None    # consumer finished reading the device buffers
        self.keys: List[str] = []


class BatchStager:
    """``depth`` slots of pinned-host + device buffers; uploads go to ``copy_stream``, consumers wait on events."""

    def __init__(self, device: torch.device | str | int = "cuda", depth: int = 2, dtypes: Optional[Mapping[str, torch.dtype]] = None):
        self.device = torch.device(device)
        if self.device.type != "cuda":
            raise PeaError("BatchStager stages into HBM; it needs a CUDA device (there is no CPU fallback)")
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.copy_stream = torch.cuda.Stream(self.device)
        self.slots = [_Slot() for _ in range(depth)]
        self.dtypes = dict(dtypes or {})   # optional per-key device dtype (the cast happens on the host copy, once)
        self._next = 0
        self.h2d_bytes = 0                 # bytes uploaded by the last stage() call

    # ---------------------------------------------------------------------------------------------- one batch
    def stage(self, host_batch: Mapping[str, Tensor]) -> int:
        """Start uploading ``host_batch``; returns the slot index to pass to :meth:`take` / :meth:`release`."""
        idx = self._next
        self._next = (self._next + 1) % len(self.slots)
        slot = self.slots[idx]
        if slot.copied is not None:
            slot.copied.synchronize()      # this slot's previous upload has left its pinned buffers
        slot.keys = list(host_batch.keys())
        nbytes = 0
        fresh = False
        for k, v in host_batch.items():
            if not isinstance(v, Tensor):
                raise TypeError(f"batch[{k!r}] is {type(v).__name__}; only tensors are staged (keep text fields on the host)")
            if v.is_cuda:
                raise PeaError(f"batch[{k!r}] is already on {v.device}; BatchStager takes host tensors")
            dt = self.dtypes.get(k, v.dtype)
            p = slot.pinned.get(k)
            if p is None or p.shape != v.shape or p.dtype != dt:
                p = slot.pinned[k] = torch.empty(v.shape, dtype=dt).pin_memory()
                slot.dev[k] = torch.empty(v.shape, dtype=dt, device=self.device)
                fresh = True
            p.copy_(v)                     # host memcpy (+ cast) into the pinned slot
            nbytes += p.numel() * p.element_size()
        self.h2d_bytes = nbytes
        if fresh:
            # a new device buffer comes from the caching allocator of the CALLER's stream and may be memory that stream's
            # pending work still uses: the first upload into it must be ordered after that work
            self.copy_stream.wait_stream(torch.cuda.current_stream(self.device))
        if slot.consumed is not None:
            self.copy_stream.wait_event(slot.consumed)   # do not overwrite device buffers a step is still reading
        with torch.cuda.stream(self.copy_stream):
            for k in slot.keys:
                slot.dev[k].copy_(slot.pinned[k], non_blocking=True)
            if slot.copied is None:
                slot.copied = torch.cuda.Event()
            slot.copied.record(self.copy_stream)
        return idx

    def take(self, idx: int) -> Dict[str, Tensor]:
        """Device tensors of slot ``idx``; the current stream waits (on the device) for the upload."""
        slot = self.slots[idx]
        if slot.copied is None:
            raise PeaError(f"slot {idx} was never staged")
        torch.cuda.current_stream(self.device).wait_event(slot.copied)
        return {k: slot.dev[k] for k in slot.keys}

    def release(self, idx: int) -> None:
        """Mark everything enqueued so far on the current stream as the last reader of slot ``idx``."""
        slot = self.slots[idx]
        if slot.consumed is None:
            slot.consumed = torch.cuda.Event()
        slot.consumed.record(torch.cuda.current_stream(self.device))

    # ---------------------------------------------------------------------------------------------- iterator form
    def prefetch(self, host_batches: Iterable[Mapping[str, Tensor]]) -> Iterator[Dict[str, Tensor]]:
        """Yield device batches while the next one uploads.  The yielded dict is valid until the following iteration."""
        it = iter(host_batches)
        try:
            cur = self.stage(next(it))
        except StopIteration:
            return
        while True:
            try:
                nxt: Optional[int] = self.stage(next(it)) if len(self.slots) > 1 else None
                exhausted = False
            except StopIteration:
                nxt, exhausted = None, True
            yield self.take(cur)
            self.release(cur)
            if exhausted:
                return
            if nxt is None:                # depth 1: nothing can overlap; upload after the consumer is done
                try:
                    nxt = self.stage(next(it))
                except StopIteration:
                    return
            cur = nxt
