#!/usr/bin/env python
"""ResidualTransformer's global attention alone: tcgen05 kernel (V^T + attention + merge) vs the mma.sync kernel.
usage: python tools/probes/attn_probe.py [B] [S]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from transformerupscaler_b200 import _lib

lib = _lib.load()
lib.tu_debug_set(b"global_attn_tc", 1)      # the tcgen05 kernel is opt-in; with a NULL workspace the mma.sync kernel runs
B = int(sys.argv[1]) if len(sys.argv) > 1 else 2
S = int(sys.argv[2]) if len(sys.argv) > 2 else 3600
heads, dim = 8, 128
torch.manual_seed(0)
qkv = (torch.randn(B * S, 3 * dim, device="cuda") * 1.0).bfloat16()
qkv[:, :dim] *= 0.25
out_new = torch.empty(B * S, dim, device="cuda", dtype=torch.bfloat16)
out_old = torch.empty_like(out_new)
n = lib.tu_global_attention_workspace_bytes(B, S, heads)
ws = torch.empty(n, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
flop = 4.0 * B * heads * S * S * 16


def run(new, iters=30):
    o, w, nb = (out_new, ws.data_ptr(), n) if new else (out_old, 0, 0)
    for _ in range(3):
        _lib.check(lib.tu_global_attention(qkv.data_ptr(), o.data_ptr(), B, S, heads, w, nb, st))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        lib.tu_global_attention(qkv.data_ptr(), o.data_ptr(), B, S, heads, w, nb, st)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


for rnd in range(3):
    a, b = run(True), run(False)
    print(f"B={B} S={S}: tcgen05 {a:.1f} us ({flop / a / 1e6:.0f} TFLOP/s)   mma.sync {b:.1f} us ({flop / b / 1e6:.0f} TFLOP/s)   "
          f"max|diff| {(out_new.float() - out_old.float()).abs().max().item():.2e}", flush=True)
