#!/usr/bin/env python
"""Copy-only ceiling of the end-to-end path: the same pinned host buffers bench.py's `e2e` loop uses
(uint8 frames: 8 x 3 x 720 x 1280 in, 8 x 3 x 1080 x 1920 out per step and rank), H2D on one stream and D2H on another, NO compute.
Run under torchrun with N ranks; rank 0 prints one JSON line: per-rank and aggregate GB/s and the frames/s the copies alone allow.

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29611 tools/probes/copy_ceiling.py
"""
import json, os, sys, time
import torch
import torch.distributed as dist

rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
numa = None
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(local)
    pynvml.nvmlDeviceSetCpuAffinity(h)
    numa = sorted(os.sched_getaffinity(0))[:4]
except Exception as e:  # noqa: BLE001
    numa = repr(e)[:60]
if world > 1:
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=dev)
FR = int(os.environ.get("TU_FRAMES", "8"))
hin = [torch.empty((FR, 3, 720, 1280), dtype=torch.uint8).pin_memory() for _ in range(2)]
hout = [torch.empty((FR, 3, 1080, 1920), dtype=torch.uint8).pin_memory() for _ in range(2)]
din = [torch.empty_like(h_, device=dev) for h_ in hin]
dout = [torch.empty_like(h_, device=dev) for h_ in hout]
s_in, s_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)


def loop(n, do_in=True, do_out=True):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(n):
        if do_in:
            with torch.cuda.stream(s_in):
                din[i & 1].copy_(hin[i & 1], non_blocking=True)
        if do_out:
            with torch.cuda.stream(s_out):
                hout[i & 1].copy_(dout[i & 1], non_blocking=True)
    s_in.synchronize(); s_out.synchronize()
    dt = time.perf_counter() - t0
    t = torch.tensor([dt], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item()


loop(10)
res = {}
n = 200
for name, a, b in (("h2d_only", True, False), ("d2h_only", False, True), ("both", True, True)):
    dt = loop(n, a, b)
    by = (hin[0].numel() if a else 0) + (hout[0].numel() if b else 0)
    res[name] = {"gbs_per_rank": by * n / dt / 1e9, "gbs_aggregate": by * n * world / dt / 1e9,
                 "frames_per_s_allowed": FR * n * world / dt}
if rank == 0:
    print(json.dumps({"n_gpus": world, "frames_per_step_per_rank": FR, "h2d_bytes": hin[0].numel(), "d2h_bytes": hout[0].numel(),
                      "cpu_affinity_head": numa, **res}), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
