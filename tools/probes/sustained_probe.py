#!/usr/bin/env python
"""Per-forward device time over a long back-to-back run (does a kernel change hold up under the power cap?).
usage: python tools/probes/sustained_probe.py [debug_key=value ...]"""
import os, sys, threading, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from transformerupscaler_b200.synth import synth_state_dict, synth_frames
from transformerupscaler_b200 import _lib
from transformerupscaler_b200.models.WindowTransformer.model import TransformerModel

lib = _lib.load()
U8 = "u8" in sys.argv[1:]
for kv in sys.argv[1:]:
    if "=" not in kv:
        continue
    k, v = kv.split("=")
    lib.tu_debug_set(k.encode(), int(v))
dev = torch.device("cuda:0")
m = TransformerModel().eval()
m.load_state_dict(synth_state_dict("WindowTransformer", 0), strict=True)
m = m.to(dev).bfloat16()
xs = [synth_frames(8, 720, 1280, seed=123 + i).to(dev).bfloat16() for i in range(2)]
if U8:      # uint8 frames in and out (the end-to-end path of bench.py)
    xs = [(x.float() * 255).round().clamp(0, 255).to(torch.uint8) for x in xs]
import pynvml
pynvml.nvmlInit()
h = pynvml.nvmlDeviceGetHandleByIndex(0)
samples, stop = [], False
def sampler():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0))
        time.sleep(0.002)
N = 120
with torch.no_grad():
    for i in range(3):
        m(xs[i & 1])
    torch.cuda.synchronize()
    th = threading.Thread(target=sampler); th.start()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(N + 1)]
    ev[0].record()
    for i in range(N):
        m(xs[i & 1])
        ev[i + 1].record()
    torch.cuda.synchronize()
    stop = True; th.join()
ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(N)]
print(sys.argv[1:], "per-forward ms:", " ".join(f"{v:.3f}" for v in ms[::6]))
print("  mean first 10 %.4f  mean last 60 %.4f  fps(last 60) %.1f" % (sum(ms[:10]) / 10, sum(ms[-60:]) / 60, 8e3 / (sum(ms[-60:]) / 60)))
ck = [s[0] for s in samples]; pw = [s[1] for s in samples]
print("  clocks MHz min/median/max %d/%d/%d  power W max %.0f  samples %d" % (min(ck), sorted(ck)[len(ck) // 2], max(ck), max(pw), len(ck)))
