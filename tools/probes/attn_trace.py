#!/usr/bin/env python
"""Phase timeline of the tcgen05 global attention (CTA 0, first warp of each softmax group): cycles between the marks
0 start | 1 S ready | 2 S in registers | 3 row max exchanged | 4 O(t-1) ready | 5 my turn on the MUFU | 6 exponentials done | 7 P published"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from transformerupscaler_b200 import _lib
lib = _lib.load()
lib.tu_debug_set(b"global_attn_tc", 1)      # needs a build with -DTU_GA_TRACE (transformerupscaler_b200/build.py NVCC_FLAGS)
B, S, heads, dim = 16, 3600, 8, 128
qkv = torch.randn(B * S, 3 * dim, device="cuda").bfloat16()
out = torch.empty(B * S, dim, device="cuda", dtype=torch.bfloat16)
n = lib.tu_global_attention_workspace_bytes(B, S, heads)
ws = torch.empty(n, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
for _ in range(2):
    lib.tu_global_attention(qkv.data_ptr(), out.data_ptr(), B, S, heads, ws.data_ptr(), n, st)
tr = torch.zeros(2 * 96 * 8, dtype=torch.int64, device="cuda")
lib.tu_debug_trace(tr.data_ptr(), 96)
lib.tu_global_attention(qkv.data_ptr(), out.data_ptr(), B, S, heads, ws.data_ptr(), n, st)
torch.cuda.synchronize()
lib.tu_debug_trace(0, 0)
t = tr.cpu().reshape(2, 96, 8)
base = int(t[0, 40, 0])
names = ["wait S", "ld S", "max+xchg", "wait O", "wait turn", "exps", "publish"]
for g in range(2):
    print(f"group {g}: tile start (rel) | " + " | ".join(names) + " | tile total")
    for i in range(40, 52):
        r = t[g, i]
        d = [int(r[k + 1] - r[k]) for k in range(7)]
        print(f"  t{i}: {int(r[0]) - base:7d} | " + " | ".join(f"{x:6d}" for x in d) + f" | {int(t[g, i + 1, 0] - r[0]):6d}")
