#!/usr/bin/env python
"""Device-resident throughput with the forwards of consecutive batches alternating between k CUDA streams (inputs resident in HBM,
no copies): does an HBM-bound kernel of one batch fill the gaps of a tensor-bound kernel of the other?
usage: python tools/probes/multistream_probe.py [frames=8] [n=60]"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from transformerupscaler_b200.synth import synth_state_dict, synth_frames
from transformerupscaler_b200.models.WindowTransformer.model import TransformerModel

frames = int(sys.argv[1]) if len(sys.argv) > 1 else 8
n = int(sys.argv[2]) if len(sys.argv) > 2 else 60
dev = torch.device("cuda:0")
m = TransformerModel().eval()
m.load_state_dict(synth_state_dict("WindowTransformer", 0), strict=True)
m = m.to(dev).bfloat16()
xs = [synth_frames(frames, 720, 1280, seed=123 + i).to(dev).bfloat16() for i in range(4)]
with torch.no_grad():
    for k in (1, 2, 3, 1, 2):
        streams = [torch.cuda.Stream(dev) for _ in range(k)]
        for i in range(3 * k):
            with torch.cuda.stream(streams[i % k]):
                m(xs[i % 4])
        torch.cuda.synchronize()
        time.sleep(0.7)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for s in streams:
            s.wait_event(e0)
        for i in range(n):
            with torch.cuda.stream(streams[i % k]):
                m(xs[i % 4])
        for s in streams:
            torch.cuda.current_stream().wait_stream(s)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(f"{k} stream(s): {ms:.4f} ms per batch of {frames}  {frames / ms * 1e3:.0f} frames/s", flush=True)
