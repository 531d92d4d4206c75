#!/usr/bin/env python
"""GPU bring-up probe for the tcgen05 conv kernel: compares it with the CUDA-core kernel on the same bf16 inputs
and times it.  Each variant runs in its own subprocess (a device-side trap poisons the CUDA context).
  python tools/tc_probe.py            # runs all variants, writes gpurun_out/tc_probe.log
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

CASES = [  # (B, H, W, stride, relu, r)
    (1, 4, 128, 1, 0, 0), (1, 8, 128, 1, 1, 0), (2, 19, 300, 1, 1, 0), (1, 37, 53, 1, 0, 0),
    (1, 16, 256, 2, 0, 0), (1, 33, 130, 2, 0, 0), (1, 13, 140, 1, 0, 2), (1, 9, 70, 1, 0, 3),
    (1, 720, 1280, 1, 1, 0), (1, 720, 1280, 2, 0, 0),
]


def child(mode):
    import numpy as np
    import torch
    from tests import gpu_helpers as G
    from transformerupscaler_b200 import _lib
    lib = _lib.load()
    lib.tu_debug_set(b"tc_base_off_mode", mode)
    dev = torch.device("cuda:0")
    rs = np.random.RandomState(0)
    for (B, H, W, s, relu, r) in CASES:
        nch = max(1, r * r)
        x = torch.from_numpy(rs.uniform(-1, 1, (B, H, W, 64)).astype(np.float32)).to(dev, torch.bfloat16)
        w = torch.from_numpy(rs.uniform(-0.06, 0.06, (nch, 9, 64, 64)).astype(np.float32)).to(dev, torch.bfloat16)
        b = torch.from_numpy(rs.uniform(-0.1, 0.1, nch * 64).astype(np.float32)).to(dev)
        lib.tu_set_bf16_tcgen05(0)
        ref = G.conv3x3_c64(x, w, b, stride=s, relu=relu, nchunk=nch, ps_r=r)
        lib.tu_set_bf16_tcgen05(1)
        out = G.conv3x3_c64(x, w, b, stride=s, relu=relu, nchunk=nch, ps_r=r)
        torch.cuda.synchronize()
        d = (out.float() - ref.float()).abs()
        print(json.dumps({"mode": mode, "case": [B, H, W, s, relu, r], "max_abs": d.max().item(), "mean_abs": d.mean().item(),
                          "ref_absmax": ref.float().abs().max().item(), "bad_frac": (d > 0.05).float().mean().item()}), flush=True)
    # timing: conv2 of the benchmark workload
    B, H, W = 8, 720, 1280
    x = torch.randn(B, H, W, 64, device=dev).to(torch.bfloat16)
    w = (torch.randn(1, 9, 64, 64, device=dev) * 0.05).to(torch.bfloat16)
    b = torch.zeros(64, device=dev)
    for tc in (1, 0):
        lib.tu_set_bf16_tcgen05(tc)
        for _ in range(2):
            G.conv3x3_c64(x, w, b, relu=1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        n = 10 if tc else 2
        e0.record()
        for _ in range(n):
            G.conv3x3_c64(x, w, b, relu=1)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / n
        print(json.dumps({"mode": mode, "timing": "conv2 8x720x1280", "tcgen05": tc, "ms": ms,
                          "tflops": 2 * 576 * 64 * B * H * W / ms / 1e9}), flush=True)


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--mode":
        child(int(sys.argv[2]))
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "tc_probe.log"), "w") as log:
        for mode in (1, 0):
            try:
                r = subprocess.run([sys.executable, os.path.abspath(__file__), "--mode", str(mode)], capture_output=True, text=True,
                                   timeout=240, cwd=ROOT)
                txt = r.stdout + ("\nSTDERR:\n" + r.stderr[-3000:] if r.returncode else "") + f"\n[mode {mode}] rc={r.returncode}\n"
            except subprocess.TimeoutExpired as e:
                txt = f"[mode {mode}] TIMEOUT\n{(e.stdout or b'').decode() if isinstance(e.stdout, bytes) else (e.stdout or '')}\n"
            log.write(txt)
            log.flush()
            print(txt)


if __name__ == "__main__":
    main()
