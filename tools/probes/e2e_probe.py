#!/usr/bin/env python
"""End-to-end throughput of FramePipeline (pinned uint8 host frames in, uint8 frames out, cfg2: 8 frames 720p -> 1080p) for several
pipeline depths and numbers of compute streams.   usage: python tools/probes/e2e_probe.py [steps]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from transformerupscaler_b200.synth import synth_state_dict, synth_frames
from transformerupscaler_b200.models.WindowTransformer.model import TransformerModel
from transformerupscaler_b200.pipeline import FramePipeline

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 40
dev = torch.device("cuda:0")
m = TransformerModel().eval()
m.load_state_dict(synth_state_dict("WindowTransformer", 0), strict=True)
m = m.to(dev).bfloat16()
B, OH, OW = 8, 1080, 1920
xs = [synth_frames(B, 720, 1280, seed=123 + i) for i in range(2)]
hin = [(x.float() * 255).round().clamp(0, 255).to(torch.uint8).pin_memory() for x in xs]
hout = [torch.empty((B, 3, OH, OW), dtype=torch.uint8).pin_memory() for _ in range(4)]
for rnd in range(2):
    for depth, cs in ((3, 2), (3, 3), (4, 2), (4, 3), (5, 3), (6, 3), (4, 4)):
        pipe = FramePipeline(m, depth=depth, device=dev, compute_streams=cs, res_out=(OH, OW))
        for i in range(3 * depth):
            pipe.submit(hin[i & 1], hout[i & 3])
        pipe.drain()
        time.sleep(0.5)
        t0 = time.perf_counter()
        for i in range(steps):
            pipe.submit(hin[i & 1], hout[i & 3])
        pipe.drain()
        dt = time.perf_counter() - t0
        print(f"depth {depth} compute streams {cs}: {B * steps / dt:8.1f} frames/s  ({dt / steps * 1e3:.3f} ms per step)", flush=True)
