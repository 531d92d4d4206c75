#!/usr/bin/env python
"""Generate golden vectors by running the UNMODIFIED reference modules (build container only).

Run once here:  python tests/golden/make_golden.py     (needs /root/reference; CPU, fp32)
Writes tests/golden/<case>.npz with the reference's pre-clamp output ``pre``
(torch.clamp patched to identity; the clamped output is checked to equal clamp(pre, 0, 1)).  Inputs and
weights are NOT stored: they are recomputed from ``oracle.weights`` (RandomState streams).
Cases are listed in tests/golden/cases.py, shared with the tests.
"""
import importlib, os, sys
import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from oracle.weights import synth_state_dict, synth_frames      # noqa: E402
from tests.golden.cases import CASES, FULLSIZE, sample_fullsize  # noqa: E402
from tests.golden._refload import ref_model                      # noqa: E402


def run_ref(model, sd, x, kw):
    M = ref_model(model)
    M.load_state_dict(sd, strict=True)
    with torch.no_grad():
        out = M(x, **kw)
        real = torch.clamp
        torch.clamp = lambda t, *a, **k: t
        try:
            pre = M(x, **kw)
        finally:
            torch.clamp = real
    return out, pre


def main_fullsize(only=None):
    """BASELINE.json's configurations at their real sizes: `python tests/golden/make_golden.py --fullsize [name ...]`"""
    torch.set_num_threads(os.cpu_count())
    for name, c in FULLSIZE.items():
        if only and name not in only:
            continue
        sd = synth_state_dict(c["model"], c["wseed"])
        B, _, H, W = c["shape"]
        x = synth_frames(B, H, W, seed=c["xseed"])
        M = ref_model(c["model"])
        M.load_state_dict(sd, strict=True)
        real = torch.clamp
        torch.clamp = lambda t, *a, **k: t
        try:
            with torch.no_grad():
                pre = torch.cat([M(x[i:i + 1], **c["kw"]) for i in range(B)], 0)      # frames are independent (checked by the small cases)
        finally:
            torch.clamp = real
        lat, crops = sample_fullsize(pre.numpy(), c)
        np.savez_compressed(os.path.join(HERE, name + ".npz"), pre=lat, shape=np.array(pre.shape),
                            **{f"crop{i}": cr for i, cr in enumerate(crops)})
        out = pre.clamp(0, 1)
        print(name, tuple(pre.shape), "mean %.4f sat0 %.3f sat1 %.3f" % (
            out.mean().item(), (out == 0).float().mean().item(), (out == 1).float().mean().item()), flush=True)


def main():
    if "--fullsize" in sys.argv:
        return main_fullsize([a for a in sys.argv[1:] if not a.startswith("--")])
    torch.set_num_threads(os.cpu_count())
    for name, c in CASES.items():
        sd = synth_state_dict(c["model"], c["wseed"], c.get("gain", 1.0))
        B, _, H, W = c["shape"]
        x = synth_frames(B, H, W, seed=c["xseed"])
        out, pre = run_ref(c["model"], sd, x, c["kw"])
        st = c.get("stride", 1)
        assert torch.equal(out, pre.clamp(0.0, 1.0))          # so only the pre-clamp tensor is stored
        np.savez_compressed(os.path.join(HERE, name + ".npz"),
                            pre=pre.numpy()[..., ::st, ::st], shape=np.array(out.shape))
        print(name, tuple(out.shape), "mean %.4f sat0 %.3f sat1 %.3f" % (
            out.mean().item(), (out == 0).float().mean().item(), (out == 1).float().mean().item()))


if __name__ == "__main__":
    main()
