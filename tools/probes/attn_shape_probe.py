#!/usr/bin/env python
"""ResidualTransformer's global attention (mma.sync kernel) alone for each CTA shape of the debug key "ga_shape".
usage: python tools/probes/attn_shape_probe.py [S]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from transformerupscaler_b200 import _lib

lib = _lib.load()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 3600
heads, dim = 8, 128
st = torch.cuda.current_stream().cuda_stream
for B in (1, 2, 4, 16):
    torch.manual_seed(0)
    qkv = (torch.randn(B * S, 3 * dim, device="cuda") * 1.0).bfloat16()
    qkv[:, :dim] *= 0.25
    flop = 4.0 * B * heads * S * S * 16
    ref = None
    nws = lib.tu_global_attention_workspace_bytes(B, S, heads)
    ws = torch.empty(max(nws, 16), dtype=torch.uint8, device="cuda")
    for rnd in range(2):
        line = []
        for shape in (0, 2, 4, 5, -1):
            lib.tu_debug_set(b"ga_shape", shape)
            out = torch.empty(B * S, dim, device="cuda", dtype=torch.bfloat16)
            for _ in range(3):
                _lib.check(lib.tu_global_attention(qkv.data_ptr(), out.data_ptr(), B, S, heads, ws.data_ptr(), nws, st))
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                lib.tu_global_attention(qkv.data_ptr(), out.data_ptr(), B, S, heads, ws.data_ptr(), nws, st)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 20 * 1e3
            if ref is None:
                ref = out.clone()
            line.append(f"shape {shape}: {us:7.1f} us {flop / us / 1e6:5.0f} TF maxdiff={(out.float() - ref.float()).abs().max().item():.1e}")
        print(f"B={B} S={S}  " + "   ".join(line), flush=True)
