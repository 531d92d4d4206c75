#!/usr/bin/env python
"""Where does a window-stack block spend its time?  Phase boundaries of math warps 0 and 15 (tu_debug_trace kinds 10 / 11,
value = tile * 256 + block * 16 + phase) inside one WindowTransformer forward of 8 frames 720p.
usage: python tools/probes/stack_phase_trace.py [key=value ...]   (debug switches applied before the traced forward)"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from transformerupscaler_b200.synth import synth_state_dict, synth_frames
from transformerupscaler_b200 import _lib
from transformerupscaler_b200.models.WindowTransformer.model import TransformerModel

PH = ["block start", "LN1 stored", "qkv[0] arrived", "qkv epilogue done", "attention stored", "proj arrived", "LN2 stored",
      "fc1 c0 arrived", "GELU c0 stored", "GELU c1 stored", "GELU c2 stored", "GELU c3 stored", "fc2 arrived"]
LAST = len(PH) - 1
PH192 = ["block start", "LN1 stored", "qkv group 0 arrived", "6 x (qkv epilogue + attention)", "proj arrived", "LN2 stored",
         "fc1 q0 arrived", "GELU q0 stored", "GELU q1 stored", "GELU q2 stored", "GELU q3 stored", "fc2 arrived"]
lib = _lib.load()
model_name, frames, kw = "WindowTransformer", 8, {}
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    if k == "model":
        model_name = v
        if v == "FastTransformer":
            from transformerupscaler_b200.models.FastTransformer.model import TransformerModel
            PH, frames, kw = PH192, 4, {"upscale_factor": 2}
            LAST = len(PH) - 1
        continue
    lib.tu_debug_set(k.encode(), int(v))
dev = torch.device("cuda:0")
m = TransformerModel().eval()
m.load_state_dict(synth_state_dict(model_name, 0), strict=True)
m = m.to(dev).bfloat16()
x = synth_frames(frames, 720, 1280, seed=123).to(dev).bfloat16()
CAP = 120000
buf = torch.zeros(1 + 2 * CAP, dtype=torch.int64, device=dev)
with torch.no_grad():
    for _ in range(4):
        m(x, **kw)
    torch.cuda.synchronize()
    lib.tu_debug_trace(buf.data_ptr(), CAP)
    m(x, **kw)
    torch.cuda.synchronize()
    lib.tu_debug_trace(0, 0)
h = buf.cpu().numpy()
n = min(int(h[0]), CAP)
print("events", int(h[0]))
for kind, name in ((10, "warp 0"), (11, "warp 15")):
    ts = {}
    for i in range(n):
        meta = int(h[2 + 2 * i])
        if (meta >> 48) != kind:
            continue
        v = meta & 0xFFFFFFFF
        ts[(v >> 8, (v >> 4) & 15, v & 15)] = int(h[1 + 2 * i])
    tiles = sorted({k[0] for k in ts})
    for sel, label in ((lambda t: t < 148, "first round (tiles < 148)"), (lambda t: t >= 148, "second round")):
        d = collections.defaultdict(list)
        blk = []
        for t in tiles:
            if not sel(t):
                continue
            for b in range(8 if model_name == "WindowTransformer" else 6):
                if (t, b, 0) in ts and (t, b, LAST) in ts:
                    blk.append(ts[(t, b, LAST)] - ts[(t, b, 0)])
                for ph in range(1, LAST + 1):
                    if (t, b, ph) in ts and (t, b, ph - 1) in ts:
                        d[ph].append(ts[(t, b, ph)] - ts[(t, b, ph - 1)])
                if b and (t, b, 0) in ts and (t, b - 1, LAST) in ts:
                    d[0].append(ts[(t, b, 0)] - ts[(t, b - 1, LAST)])
        if not blk:
            continue
        print("%s, %s: block median %.2f us (n %d)" % (name, label, np.median(blk) / 1e3, len(blk)))
        for ph in range(0, LAST + 1):
            if d[ph]:
                a = np.array(d[ph]) / 1e3
                print("   -> %-20s median %6.2f us   p10 %6.2f  p90 %6.2f" % (PH[ph], np.median(a), np.percentile(a, 10), np.percentile(a, 90)))
