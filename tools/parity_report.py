#!/usr/bin/env python
"""Measured parity of the engine against the reference-generated goldens at BASELINE.json's real sizes (tests/golden/cases.py FULLSIZE):
max-abs error of the fp32 path and error of the bf16 path, pre-clamp and clamped, on the stored lattice + corner crops.
Writes gpurun_out/parity_<tag>.json (the same comparisons tests/test_gpu_models.py asserts on)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
from tests.golden.cases import FULLSIZE, sample_fullsize
from tests.test_gpu_models import build, engine_pre_clamp, bf16_pre_clamp_err, psnr
from transformerupscaler_b200.synth import synth_frames

tag = sys.argv[1] if len(sys.argv) > 1 else "x"
out = []
for name, c in FULLSIZE.items():
    B, _, H, W = c["shape"]
    M, sd = build(c["model"], c["wseed"])
    x = synth_frames(B, H, W, seed=c["xseed"]).cuda()
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    rec = {"case": name, "model": c["model"], "shape": list(c["shape"]), "kw": {k: list(v) if isinstance(v, tuple) else v for k, v in c["kw"].items()}}
    for bf16 in (False, True):
        pre = engine_pre_clamp(M, x, c["kw"], bf16=bf16).cpu().numpy()
        lat, crops = sample_fullsize(pre, c)
        refs = [g["pre"]] + [g[f"crop{i}"] for i in range(len(crops))]
        ours = [lat] + crops
        key = "bf16" if bf16 else "fp32"
        rec[key + "_pre_clamp_max_abs"] = float(max(np.abs(a - b).max() for a, b in zip(ours, refs)))
        rec[key + "_pre_clamp_rel_above_1"] = float(max(bf16_pre_clamp_err(a, b) for a, b in zip(ours, refs)))
        rec[key + "_clamped_max_abs"] = float(max(np.abs(np.clip(a, 0, 1) - np.clip(b, 0, 1)).max() for a, b in zip(ours, refs)))
        rec[key + "_clamped_psnr_db"] = psnr(torch.from_numpy(np.clip(lat, 0, 1)), torch.from_numpy(np.clip(g["pre"], 0, 1)))
        rec["ref_abs_max"] = float(np.abs(g["pre"]).max())
        del pre
        torch.cuda.empty_cache()
    print(json.dumps(rec), flush=True)
    out.append(rec)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"parity_{tag}.json"), "w"), indent=1)
