#!/usr/bin/env python
"""Timeline of the window stack / unembed overlap inside one forward (tu_debug_trace): when does each stack tile begin and end,
when does each unembed CTA draw its tiles, how long does it wait for a tile to be published.
usage: python tools/probes/overlap_trace.py [key=value ...]   (debug switches applied before the traced forward)"""
import os, sys, collections
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from transformerupscaler_b200.synth import synth_state_dict, synth_frames
from transformerupscaler_b200 import _lib
from transformerupscaler_b200.models.WindowTransformer.model import TransformerModel

lib = _lib.load()
for kv in sys.argv[1:]:
    k, v = kv.split("=")
    lib.tu_debug_set(k.encode(), int(v))
dev = torch.device("cuda:0")
m = TransformerModel().eval()
m.load_state_dict(synth_state_dict("WindowTransformer", 0), strict=True)
m = m.to(dev).bfloat16()
x = synth_frames(8, 720, 1280, seed=123).to(dev).bfloat16()
CAP = 40000
buf = torch.zeros(1 + 2 * CAP, dtype=torch.int64, device=dev)
with torch.no_grad():
    for _ in range(4):
        m(x)
    torch.cuda.synchronize()
    lib.tu_debug_trace(buf.data_ptr(), CAP)
    m(x)
    torch.cuda.synchronize()
    lib.tu_debug_trace(0, 0)
h = buf.cpu().numpy()
n = int(h[0])
ev = [(int(h[1 + 2 * i]), int(h[2 + 2 * i]) >> 48, (int(h[2 + 2 * i]) >> 32) & 0xFFFF, int(h[2 + 2 * i]) & 0xFFFFFFFF) for i in range(min(n, CAP))]
t0 = min(e[0] for e in ev)
ev = sorted((t - t0, k, sm, v) for t, k, sm, v in ev)
print("events", n)
begin = {v: t for t, k, sm, v in ev if k == 1}
end = {v: t for t, k, sm, v in ev if k == 2}
sm_of = {v: sm for t, k, sm, v in ev if k == 1}
ends = sorted(end.values())
print("stack: first tile begins %.1f us, tile ends (us): min %.1f  p25 %.1f  median %.1f  p75 %.1f  max %.1f" % (
    min(begin.values()) / 1e3, ends[0] / 1e3, ends[len(ends) // 4] / 1e3, ends[len(ends) // 2] / 1e3, ends[3 * len(ends) // 4] / 1e3, ends[-1] / 1e3))
dur = sorted((end[v] - begin[v]) / 1e3 for v in end)
print("stack: tile duration us: min %.1f median %.1f max %.1f;  tiles %d on %d SMs" % (dur[0], dur[len(dur) // 2], dur[-1], len(end), len(set(sm_of.values()))))
first_round = [v for v in end if begin[v] < 20e3]
second = [v for v in end if begin[v] >= 20e3]
print("stack: %d tiles begin in the first 20 us, %d later (begin median %.1f us)" % (len(first_round), len(second), sorted(begin[v] for v in second)[len(second) // 2] / 1e3 if second else -1))
draws = [(t, sm, v) for t, k, sm, v in ev if k == 3]
by_sm = collections.defaultdict(list)
for t, sm, v in draws:
    by_sm[sm].append((t, v))
starts = sorted(min(t for t, v in lst) for lst in by_sm.values())
print("unembed: %d CTAs drew tiles; first draw per CTA (us): min %.1f  p25 %.1f median %.1f p75 %.1f max %.1f" % (
    len(by_sm), starts[0] / 1e3, starts[len(starts) // 4] / 1e3, starts[len(starts) // 2] / 1e3, starts[3 * len(starts) // 4] / 1e3, starts[-1] / 1e3))
dt = sorted(t for t, sm, v in draws)
tot = len(dt)
for frac in (0.1, 0.25, 0.5, 0.75, 0.9, 1.0):
    print("unembed: %3d %% of the %d tile draws done at %.1f us" % (int(frac * 100), tot, dt[min(int(frac * tot), tot - 1)] / 1e3))
w4 = {(sm, v): t for t, k, sm, v in ev if k == 4}
w5 = {(sm, v): t for t, k, sm, v in ev if k == 5}
waits = sorted((w5[key] - w4[key]) / 1e3 for key in w5 if key in w4)
if waits:
    print("unembed: waits for a published M tile: n %d  total %.1f us  median %.2f  max %.1f" % (len(waits), sum(waits), waits[len(waits) // 2], waits[-1]))
print("last event at %.1f us" % (ev[-1][0] / 1e3))
