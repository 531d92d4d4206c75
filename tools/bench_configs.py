#!/usr/bin/env python
"""Secondary measurements (not the headline bench): the other BASELINE.json configs through the public model classes,
device-resident, CUDA events, with the per-kernel breakdown of tu_profile_*.  Writes gpurun_out/configs_<tag>.json.

  python tools/bench_configs.py [tag] [--only substring[,substring...]]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29650 tools/bench_configs.py r2 --only cfg4_fast_720p_x2,cfg5a
      (frame-sharded: every rank runs the case's batch on its own GPU -- B frames per GPU, no data-path collective; the timing is the
       max over ranks after a barrier, fps = world * B / that)
"""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

from transformerupscaler_b200.synth import synth_state_dict, synth_frames  # noqa: E402
from transformerupscaler_b200 import _lib  # noqa: E402

# GFLOP per frame of the reference op graph
CASES = [
    ("cfg1_fast_360x640_x2_fp32", "FastTransformer", (1, 360, 640), dict(upscale_factor=2), "fp32", 139.84),
    ("cfg1_fast_360x640_x2_bf16", "FastTransformer", (1, 360, 640), dict(upscale_factor=2), "bf16", 139.84),
    ("cfg2_window_720p_1080p_b8", "WindowTransformer", (8, 720, 1280), dict(res_out=(1080, 1920)), "bf16", 126.94),
    ("cfg2_window_720p_1080p_b1", "WindowTransformer", (1, 720, 1280), dict(res_out=(1080, 1920)), "bf16", 126.94),
    ("cfg4_fast_720p_x2_b4", "FastTransformer", (4, 720, 1280), dict(upscale_factor=2), "bf16", 559.36),
    ("cfg4_fast_720p_x3_b4", "FastTransformer", (4, 720, 1280), dict(upscale_factor=3), "bf16", 916.51),
    ("cfg4_fast_720p_x4_b2", "FastTransformer", (2, 720, 1280), dict(upscale_factor=4), "bf16", 1688.92),
    ("cfg4_fast_720p_x6_b2", "FastTransformer", (2, 720, 1280), dict(upscale_factor=6), "bf16", 2845.16),
    ("cfg5a_residual_720p_4k_b2", "ResidualTransformer", (2, 720, 1280), dict(res_out=(2160, 3840)), "bf16", 179.45),
    ("cfg5b_fast_1080p_x2_b2", "FastTransformer", (2, 1080, 1920), dict(upscale_factor=2), "bf16", 1247.79),
    ("fast_720p_res1080p_b2", "FastTransformer", (2, 720, 1280), dict(res_out=(1080, 1920)), "bf16", 559.36),
]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 and not sys.argv[1].startswith("--") else "x"
    only = sys.argv[sys.argv.index("--only") + 1] if "--only" in sys.argv else ""
    import importlib
    import torch.distributed as dist
    lib = _lib.load()
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)                 # NCCL's banner goes to stderr
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)
    out = []
    for name, model, (B, H, W), kw, prec, gflop in CASES:
        if only and not any(o in name for o in only.split(",")):
            continue
        M = importlib.import_module(f"transformerupscaler_b200.models.{model}.model").TransformerModel().eval()
        M.load_state_dict(synth_state_dict(model, 0), strict=True)
        M = M.to(dev)
        x = synth_frames(B, H, W, seed=5 + rank).to(dev)
        if prec == "bf16":
            M, x = M.bfloat16(), x.bfloat16()
        rec = {"case": name, "model": model, "shape": [B, 3, H, W], "kw": {k: list(v) if isinstance(v, tuple) else v for k, v in kw.items()},
               "precision": prec}
        try:
            with torch.no_grad():
                for _ in range(3):
                    y = M(x, **kw)
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                iters = 10
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(iters):
                    y = M(x, **kw)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / iters
                if world > 1:
                    tms = torch.tensor([ms], device=dev, dtype=torch.float64)
                    dist.all_reduce(tms, op=dist.ReduceOp.MAX)
                    ms = tms.item()
                lib.tu_profile_reset()
                lib.tu_profile_enable(2)
                for _ in range(3):
                    y = M(x, **kw)
                torch.cuda.synchronize()
                lib.tu_profile_enable(0)
                n = lib.tu_profile_report(None, 0)
                buf = C.create_string_buffer(max(n, 16))
                lib.tu_profile_report(buf, len(buf))
                br = {}
                for ln in buf.value.decode().splitlines():
                    k, tot, cnt = ln.split()
                    br[k] = round(float(tot) / 3, 4)            # ms per forward (all launches of that op)
                lib.tu_profile_reset()
            rec.update(n_gpus=world, frames_per_gpu=B, ms_per_batch=round(ms, 4), fps=round(world * B / ms * 1e3, 1),
                       model_tflops_per_gpu=round(gflop * B / ms, 1),
                       flops_note="reference-graph FLOPs (SURVEY.md 8d) over the measured time: FastTransformer's folded up1 branch executes 8.04x fewer",
                       out_shape=list(y.shape), ms_by_op=br, peak_mem_gb=round(torch.cuda.max_memory_allocated() / 2**30, 2))
        except Exception as ex:  # noqa: BLE001
            rec["error"] = repr(ex)[:300]
        if rank == 0:
            print(json.dumps(rec), flush=True)
        out.append(rec)
        del M, x
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
    if rank == 0:
        os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
        json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"configs_{tag}{'_n%d' % world if world > 1 else ''}.json"), "w"), indent=1)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
