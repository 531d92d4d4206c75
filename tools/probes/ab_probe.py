#!/usr/bin/env python
"""A/B of debug switches in one process: whole-forward device time (CUDA events over N back-to-back forwards) and the
per-kernel breakdown (tu_profile mode 2) for each setting, interleaved over several rounds so that clock / power drift
hits every setting alike.
usage: python tools/probes/ab_probe.py key=v0,v1,... [model=WindowTransformer] [frames=8] [n=40] [rounds=3]"""
import os, sys, time, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from transformerupscaler_b200.synth import synth_state_dict, synth_frames
from transformerupscaler_b200 import _lib

lib = _lib.load()
opts = dict(kv.split("=") for kv in sys.argv[1:])
model_name = opts.pop("model", "WindowTransformer")
frames = int(opts.pop("frames", 8))
N = int(opts.pop("n", 40))
rounds = int(opts.pop("rounds", 3))
cool = float(opts.pop("cool", 0))      # seconds of idle before every measurement (and no long warm-up): the regime of a short bench.py run
kw = {}
if "scale" in opts:
    kw["upscale_factor"] = int(opts.pop("scale"))
(key, vals), = opts.items()
vals = [int(v) for v in vals.split(",")]
import importlib
TransformerModel = importlib.import_module(f"transformerupscaler_b200.models.{model_name}.model").TransformerModel
dev = torch.device("cuda:0")
m = TransformerModel().eval()
m.load_state_dict(synth_state_dict(model_name, 0), strict=True)
m = m.to(dev).bfloat16()
xs = [synth_frames(frames, 720, 1280, seed=123 + i).to(dev).bfloat16() for i in range(2)]


def breakdown():
    lib.tu_profile_reset()
    lib.tu_profile_enable(2)
    for i in range(10):
        m(xs[i & 1], **kw)
    torch.cuda.synchronize()
    lib.tu_profile_enable(0)
    nbuf = lib.tu_profile_report(None, 0)
    buf = C.create_string_buffer(max(nbuf, 16))
    lib.tu_profile_report(buf, len(buf))
    out = {}
    for ln in buf.value.decode().splitlines():
        name, tot, cnt = ln.split()
        out[name] = float(tot) / max(int(cnt), 1)
    lib.tu_profile_reset()
    return out


res = {v: [] for v in vals}
ref = None
with torch.no_grad():
    for v in vals:
        lib.tu_debug_set(key.encode(), v)
        y = m(xs[0], **kw).float()
        if ref is None:
            ref = y
        else:
            print(f"{key}={v}: max-abs vs {key}={vals[0]}: {(y - ref).abs().max().item():.3e}")
    for i in range(0 if cool else int(os.environ.get("AB_WARM", "150"))):      # reach the board's steady state (power cap) before comparing
        m(xs[i & 1], **kw)
    torch.cuda.synchronize()
    for r in range(rounds):
        for v in vals:
            lib.tu_debug_set(key.encode(), v)
            if cool:
                time.sleep(cool)
            for i in range(3):
                m(xs[i & 1], **kw)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(N):
                m(xs[i & 1], **kw)
            e1.record()
            torch.cuda.synchronize()
            res[v].append(e0.elapsed_time(e1) / N)
    for v in vals:
        lib.tu_debug_set(key.encode(), v)
        bd = breakdown()
        ms = res[v]
        med = sorted(ms)[len(ms) // 2]
        print(f"{key}={v}: forward ms " + " ".join(f"{t:.4f}" for t in ms) + f"  best {min(ms):.4f}  median {med:.4f}  fps(median) {frames * 1e3 / med:.1f}")
        print("    " + "  ".join(f"{k} {t * 1e3:.1f}" for k, t in bd.items()))
