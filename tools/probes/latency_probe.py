#!/usr/bin/env python
"""Batch-1 latency (the reference's speed_test.py / app_overlay.py call pattern): eager forward vs CUDA-graph replay (graph.GraphedModel).
usage: python tools/probes/latency_probe.py [model] [B]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import importlib
import torch
from transformerupscaler_b200.synth import synth_state_dict, synth_frames
from transformerupscaler_b200.graph import GraphedModel

model = sys.argv[1] if len(sys.argv) > 1 else "WindowTransformer"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1
kw = dict(upscale_factor=2) if model == "FastTransformer" else dict(res_out=(1080, 1920))
M = importlib.import_module(f"transformerupscaler_b200.models.{model}.model").TransformerModel().eval()
M.load_state_dict(synth_state_dict(model, 0), strict=True)
M = M.to("cuda:0").bfloat16()
x = synth_frames(B, 720, 1280, seed=1).cuda().bfloat16()
G = GraphedModel(M)


def bench(fn, n=200):
    with torch.no_grad():
        for _ in range(10):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        host_issue = time.perf_counter() - t0
        torch.cuda.synchronize()
        wall = time.perf_counter() - t0
        # synchronous latency: one call at a time, result waited for (what a frame-by-frame caller sees)
        t1 = time.perf_counter()
        for _ in range(50):
            fn()
            torch.cuda.synchronize()
        sync_lat = (time.perf_counter() - t1) / 50
    return e0.elapsed_time(e1) / n, host_issue / n * 1e3, wall / n * 1e3, sync_lat * 1e3


for name, fn in (("eager", lambda: M(x, **kw)), ("cuda graph", lambda: G(x, **kw))):
    dev_ms, host_ms, wall_ms, lat_ms = bench(fn)
    print(f"{model} B={B} {name:10s}: device {dev_ms:.3f} ms/call   host issue {host_ms:.3f} ms/call   wall {wall_ms:.3f} ms/call   "
          f"synchronous latency {lat_ms:.3f} ms", flush=True)
