#!/usr/bin/env python
"""Which way of putting an NCCL all-reduce into a CUDA-graph capture works with this torch / NCCL pair?  One variant per run
(a hang must not block the others); launch with torchrun on >= 2 GPUs under `timeout`:

    timeout 100 torchrun --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/nccl_capture_probe.py V

    sync     all_reduce(async_op=False) on the capturing stream
    async    all_reduce(async_op=True) on a forked side stream, work.wait() + join inside the capture
    bwd      the `async` launch issued from inside an autograd backward (the engine's worker thread) under capture
    bwdmain  backward captured, the collective issued from the main thread after backward() returns, still inside the capture
"""
import os
import sys
import threading
import time

import torch
import torch.distributed as dist

variant = sys.argv[1]
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
threading.Thread(target=lambda: (time.sleep(80), print(f"[{variant}] rank {rank}: WATCHDOG exit", flush=True), os._exit(3)), daemon=True).start()
dist.init_process_group("nccl", device_id=dev)
mode = os.environ.get("PROBE_MODE", "thread_local")

w = torch.randn(1024, 1024, device=dev, requires_grad=True)
x = torch.randn(256, 1024, device=dev)
buf = torch.zeros(1 << 20, device=dev)
comm = torch.cuda.Stream()
pending = []


def launch_async():
    comm.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(comm):
        pending.append(dist.all_reduce(buf, op=dist.ReduceOp.AVG, async_op=True))


def join():
    for wk in pending:
        wk.wait()
    pending.clear()
    torch.cuda.current_stream().wait_stream(comm)


class Sink(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a):
        return a.clone()

    @staticmethod
    def backward(ctx, g):
        buf[: g.numel()].copy_(g.reshape(-1))
        if variant == "bwd":
            launch_async()
        return g


def body():
    if variant == "sync":
        buf.view(4, -1).copy_((x @ w).detach().reshape(1, -1).expand(4, -1))
        dist.all_reduce(buf, op=dist.ReduceOp.AVG)
    elif variant == "async":
        buf.view(4, -1).copy_((x @ w).detach().reshape(1, -1).expand(4, -1))
        launch_async()
        join()
    else:
        w.grad = None
        y = Sink.apply(x @ w)
        y.square().sum().backward()
        if variant == "bwdmain":
            launch_async()
        join()
    return buf.sum()


side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    for _ in range(3):
        body()
torch.cuda.current_stream().wait_stream(side)
torch.cuda.synchronize()
dist.barrier()
print(f"[{variant}] rank {rank}: eager ok", flush=True)
g = torch.cuda.CUDAGraph()
try:
    with torch.cuda.graph(g, stream=side, capture_error_mode=mode):
        out = body()
    print(f"[{variant}] rank {rank}: capture ok (mode {mode})", flush=True)
    for _ in range(5):
        g.replay()
    torch.cuda.synchronize()
    ref = body()
    torch.cuda.synchronize()
    print(f"[{variant}] rank {rank}: replay ok, out {float(out):.6e} eager {float(ref):.6e}", flush=True)
except Exception as e:  # noqa: BLE001
    print(f"[{variant}] rank {rank}: FAILED {type(e).__name__}: {str(e).splitlines()[0][:200]}", flush=True)
    os._exit(1)
dist.barrier()
dist.destroy_process_group()
