"""Deterministic synthetic weights and frames for the three upscaler models (benchmarks, smoke test, test fixtures).

Pure generators: no arithmetic of the forward pass lives here.  (`oracle/weights.py` re-exports these names for the tests.)


``synth_state_dict(model, seed)`` returns a state_dict with exactly the reference's
keys and shapes (WindowTransformer/model.py:187-222,
FastTransformer/model.py:189-229 + utils.py:43-98, ResidualTransformer/model.py:69-112),
filled from ``numpy.random.RandomState`` (a frozen, version-stable stream) so the golden
generator (which loads them into the *reference* modules) and the tests (which load them
into the engine / the oracle) see identical bits without shipping megabytes of weights.
Magnitudes follow PyTorch's default inits (uniform +-1/sqrt(fan_in) for conv / linear,
N(0, 0.02) bias table, N(0,1) pos_embed); LayerNorm affine is perturbed away from (1, 0)
so that a missing gamma/beta would be caught.
"""
from __future__ import annotations

import math
from collections import OrderedDict

import numpy as np
import torch


def _rel_index(ws: int = 8) -> torch.Tensor:
    ys, xs = np.meshgrid(np.arange(ws), np.arange(ws), indexing="ij")
    ys, xs = ys.reshape(-1), xs.reshape(-1)
    idx = (ys[:, None] - ys[None, :] + ws - 1) * (2 * ws - 1) + (xs[:, None] - xs[None, :] + ws - 1)
    return torch.from_numpy(idx.astype(np.int64))


def _spec(model: str):
    """Ordered (name, shape, kind, fan_in) list. kind: u=uniform(+-1/sqrt(fan)), ln_w, ln_b, tbl, pos, idx."""
    s = []

    def conv(name, co, ci, k, bias=True):
        s.append((name + ".weight", (co, ci, k, k), "u", ci * k * k))
        if bias:
            s.append((name + ".bias", (co,), "u", ci * k * k))

    def lin(name, o, i):
        s.append((name + ".weight", (o, i), "u", i))
        s.append((name + ".bias", (o,), "u", i))

    def ln(name, d):
        s.append((name + ".weight", (d,), "ln_w", 0))
        s.append((name + ".bias", (d,), "ln_b", 0))

    if model == "WindowTransformer":
        dim, heads, nb = 128, 8, 8
    elif model == "FastTransformer":
        dim, heads, nb = 192, 12, 6
    elif model == "ResidualTransformer":
        dim, heads, nb = 128, 8, 8
    else:
        raise KeyError(model)

    conv("conv1", 64, 3, 3)
    conv("conv2", 64, 64, 3)
    if model == "FastTransformer":
        for pre, n in (("up1", 64), ("final_upscale", 3)):
            conv(f"{pre}.upsamplers.2.0", 4 * n, n, 3)
            conv(f"{pre}.upsamplers.3.0", 9 * n, n, 3)
            conv(f"{pre}.upsamplers.4.0", 4 * n, n, 3)
            conv(f"{pre}.upsamplers.4.2", 4 * n, n, 3)
            conv(f"{pre}.upsamplers.6.0", 36 * n, n, 3)
            if pre == "up1":
                conv("up1_conv.conv", 3, 64, 3, bias=False)
        conv("final_upscale_conv", 3, 3, 3)
    else:
        conv("downsample", 64, 64, 3)
    conv("patch_embed", dim, 64, 8)
    if model == "ResidualTransformer":
        s.append(("pos_embed", (1, 3600, dim), "pos", 0))
        for i in range(nb):
            p = f"transformer_blocks.{i}."
            ln(p + "norm1", dim)
            s.append((p + "attn.in_proj_weight", (3 * dim, dim), "u", dim))
            s.append((p + "attn.in_proj_bias", (3 * dim,), "u", dim))
            lin(p + "attn.out_proj", dim, dim)
            ln(p + "norm2", dim)
            lin(p + "mlp.0", 4 * dim, dim)
            lin(p + "mlp.2", dim, 4 * dim)
    else:
        for i in range(nb):
            p = f"window_blocks.{i}."
            ln(p + "norm1", dim)
            s.append((p + "attn.relative_position_bias_table", (225, heads), "tbl", 0))
            s.append((p + "attn.relative_position_index", (64, 64), "idx", 0))
            lin(p + "attn.qkv", 3 * dim, dim)
            lin(p + "attn.proj", dim, dim)
            ln(p + "norm2", dim)
            lin(p + "mlp.0", 4 * dim, dim)
            lin(p + "mlp.2", dim, 4 * dim)
    # ConvTranspose2d(dim, 64, 8, 8): weight (dim, 64, 8, 8); torch's fan_in for it is 64*8*8
    s.append(("patch_unembed.weight", (dim, 64, 8, 8), "u", 64 * 64))
    s.append(("patch_unembed.bias", (64,), "u", 64 * 64))
    conv("decoder_conv1", 64, 64, 3)
    conv("decoder_conv2", 3, 64, 3)
    return s


def synth_state_dict(model: str, seed: int = 0, gain: float = 1.0) -> "OrderedDict[str, torch.Tensor]":
    rs = np.random.RandomState(seed)
    sd = OrderedDict()
    for name, shape, kind, fan in _spec(model):
        if kind == "u":
            bound = gain / math.sqrt(fan)
            a = rs.uniform(-bound, bound, size=shape)
        elif kind == "ln_w":
            a = 1.0 + 0.1 * rs.standard_normal(size=shape)
        elif kind == "ln_b":
            a = 0.05 * rs.standard_normal(size=shape)
        elif kind == "tbl":
            a = 0.02 * rs.standard_normal(size=shape)
        elif kind == "pos":
            a = rs.standard_normal(size=shape)
        elif kind == "idx":
            sd[name] = _rel_index(8)
            continue
        sd[name] = torch.from_numpy(np.asarray(a, dtype=np.float32))
    return sd


def synth_frames(batch: int, h: int, w: int, seed: int = 123) -> torch.Tensor:
    """Synthetic RGB frames in [0,1): smooth low-frequency content plus noise, fp32 NCHW."""
    rs = np.random.RandomState(seed)
    yy, xx = np.meshgrid(np.linspace(0, 1, h), np.linspace(0, 1, w), indexing="ij")
    out = np.empty((batch, 3, h, w), np.float32)
    for b in range(batch):
        for c in range(3):
            fx, fy, ph = rs.uniform(1, 6), rs.uniform(1, 6), rs.uniform(0, 6.28)
            base = 0.5 + 0.35 * np.sin(6.28 * (fx * xx + fy * yy) + ph)
            out[b, c] = np.clip(base + 0.15 * rs.uniform(-1, 1, size=(h, w)), 0, 0.999)
    return torch.from_numpy(out)
