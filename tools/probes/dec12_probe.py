#!/usr/bin/env python
"""Decoder convolutions alone, through the C ABI: the fused kernel (both ring configurations) against the two-kernel path
(streaming 64 -> 64 + tile 64 -> 3), CUDA events over N back-to-back calls, interleaved over several rounds.
usage: python tools/probes/dec12_probe.py [B H W] [n=20] [rounds=5]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from transformerupscaler_b200 import _lib

lib = _lib.load()
args = [a for a in sys.argv[1:] if "=" not in a]
opts = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
B, H, W = (int(v) for v in args[:3]) if len(args) >= 3 else (8, 360, 640)
N, rounds = int(opts.get("n", 20)), int(opts.get("rounds", 5))
dev = torch.device("cuda:0")
rs = np.random.RandomState(3)
BF16 = torch.bfloat16
x = [torch.from_numpy(rs.uniform(-1, 1, (B, H, W, 64)).astype(np.float32)).to(dev, BF16) for _ in range(2)]
w1 = torch.from_numpy(rs.uniform(-0.05, 0.05, (64, 64, 3, 3)).astype(np.float32))
w2 = torch.from_numpy(rs.uniform(-0.05, 0.05, (3, 64, 3, 3)).astype(np.float32))
b1 = torch.from_numpy(rs.uniform(-0.1, 0.1, 64).astype(np.float32)).to(dev)
b2 = torch.from_numpy(rs.uniform(-0.1, 0.1, 3).astype(np.float32)).to(dev)
w1p = w1.permute(2, 3, 0, 1).reshape(9, 64, 64).contiguous().to(dev, BF16)
w16 = torch.zeros(3, 4, 4, 64)
w16[:, :3, :3] = w2.permute(2, 3, 0, 1)
w16 = w16.reshape(3, 16, 64).to(dev, BF16)
w2p = w2.permute(2, 3, 1, 0).reshape(9, 64, 3).contiguous().to(dev)
mid = torch.empty(B, H, W, 64, dtype=BF16, device=dev)
out = torch.empty(B, 3, H, W, dtype=torch.float32, device=dev)
st = lambda: torch.cuda.current_stream().cuda_stream
p = lambda t: t.data_ptr()


def two(i):
    _lib.check(lib.tu_conv3x3_c64(p(x[i & 1]), p(w1p), p(b1), p(mid), 1, B, H, W, 1, 1, 1, 0, st()))
    _lib.check(lib.tu_conv3x3_c64_to3(p(mid), 1, p(w2p), p(w16), p(b2), p(out), B, H, W, 0, st()))


def fused(i):
    _lib.check(lib.tu_dec12_fused(p(x[i & 1]), p(w1p), p(b1), p(w16), p(b2), p(out), B, H, W, st()))


def run(fn):
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(N):
        fn(i)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / N * 1e3


cases = [("two kernels", None, two), ("fused 5/3/2", 1, fused), ("fused 6/2/1", 3, fused), ("5/3/2 nodefer", 5, fused), ("6/2/1 nodefer", 7, fused)]
res = {c[0]: [] for c in cases}
two(0)
ref = out.clone()
for r in range(rounds):
    for name, key, fn in cases:
        if key is not None:
            lib.tu_debug_set(b"fuse_dec12", key)
        res[name].append(run(fn))
        if r == 0 and key is not None:
            fn(0)
            torch.cuda.synchronize()
            print(f"{name}: max-abs vs two kernels {(out - ref).abs().max().item():.3e}")
lib.tu_debug_set(b"fuse_dec12", 1)
flop = 2.0 * B * H * W * 9 * 64 * (64 + 3)
for name, _, _ in cases:
    v = sorted(res[name])
    print(f"{name:14s} us: " + " ".join(f"{t:.1f}" for t in res[name]) + f"   median {v[len(v) // 2]:.1f}  ({flop / v[len(v) // 2] / 1e6:.0f} TFLOP/s)")
