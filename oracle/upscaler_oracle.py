"""TEST INFRASTRUCTURE ONLY — CPU restatement of the TransformerUpscaler forward pass.

This file is the *oracle* for the B200 engine: a plain CPU restatement of what the
reference's three ``TransformerModel.forward`` methods compute, written in the
engine's own op decomposition (3x3 convs as nine shifted GEMMs over NHWC,
patch-embed / patch-unembed as GEMM + gather/scatter, dense relative-position
bias, explicit Keys-cubic resampling, explicit PixelShuffle index map).
torch-on-CPU is used only as the array / BLAS library (``matmul``, elementwise,
``erf``); no ``nn.Module``, ``F.conv2d``, ``F.interpolate`` or ``nn.PixelShuffle``
is called, so that every semantic choice of the reference is restated here and
pinned by the golden vectors.

Reference lines restated (paths relative to the reference root):
  * WindowTransformer/model.py:224-305  -> :func:`window_forward`
  * FastTransformer/model.py:231-327    -> :func:`fast_forward`
  * ResidualTransformer/model.py:114-165 -> :func:`residual_forward`
  * WindowTransformer/model.py:29-61 (window_partition / window_reverse)
  * WindowTransformer/model.py:63-131 (WindowAttention), :133-170 (block)
  * FastTransformer/utils.py:43-98 (Upsampler = conv + PixelShuffle), :13-40 (BasicConv)
  * ResidualTransformer/model.py:22-50 (nn.MultiheadAttention block)
The arithmetic itself lives in third-party PyTorch (torch~=2.6 pinned by the
reference's requirements.txt; 2.11 in this image): ATen ``upsample_bicubic2d``
(``ATen/native/UpSample.h:259-312,400-448``: scale = in/out, src = fma(scale,
dst+0.5, -0.5), Keys cubic A=-0.75, border-clamped taps), ``nn.PixelShuffle``,
``nn.LayerNorm`` (eps 1e-5, biased variance), ``nn.GELU`` (erf form),
``nn.MultiheadAttention`` (in_proj rows [q;k;v], heads = contiguous 16-wide
chunks, scale 1/sqrt(hd)).

Pinning: the reference ships no tests or golden vectors for this path (SURVEY.md
§4, §8c), so the oracle is pinned against *outputs of the reference itself run in
the build container*: ``tests/golden/make_golden.py`` imports the reference
modules from /root/reference, loads :func:`oracle.weights.synth_state_dict`
weights, and stores input seeds + outputs under ``tests/golden/``;
``tests/test_oracle_golden.py`` checks this file against them.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import torch

Tensor = torch.Tensor
SD = Dict[str, Tensor]


# --------------------------------------------------------------------------- helpers
def _p(sd: SD, name: str, dtype) -> Tensor:
    return sd[name].detach().to("cpu", dtype)


def conv3x3_nhwc(x: Tensor, w: Tensor, b: Optional[Tensor], stride: int = 1, relu: bool = False) -> Tensor:
    """3x3, padding 1 convolution on NHWC data as nine shifted GEMMs.

    x (B,H,W,Cin), w (Cout,Cin,3,3) [nn.Conv2d layout], b (Cout) or None.
    Output size floor((H+2-3)/stride)+1 (= ceil(H/stride) for stride 2).
    """
    B, H, W, Cin = x.shape
    Ho = (H + 2 - 3) // stride + 1
    Wo = (W + 2 - 3) // stride + 1
    xp = torch.zeros(B, H + 2, W + 2, Cin, dtype=x.dtype)
    xp[:, 1:H + 1, 1:W + 1] = x
    out = torch.zeros(B, Ho, Wo, w.shape[0], dtype=x.dtype)
    for ky in range(3):
        for kx in range(3):
            a = xp[:, ky:ky + stride * (Ho - 1) + 1:stride, kx:kx + stride * (Wo - 1) + 1:stride, :]
            out += a.reshape(-1, Cin).matmul(w[:, :, ky, kx].t()).reshape(B, Ho, Wo, -1)
    if b is not None:
        out += b
    if relu:
        out.clamp_(min=0)
    return out


def pixel_shuffle_nhwc(x: Tensor, r: int) -> Tensor:
    """nn.PixelShuffle on NHWC: out[b, h*r+i, w*r+j, c] = in[b, h, w, c*r*r + i*r + j]."""
    B, H, W, C = x.shape
    c = C // (r * r)
    x = x.reshape(B, H, W, c, r, r)          # (b,h,w,c,i,j)
    x = x.permute(0, 1, 4, 2, 5, 3)          # (b,h,i,w,j,c)
    return x.reshape(B, H * r, W * r, c).contiguous()


def _cubic_taps(in_size: int, out_size: int, dtype) -> Tuple[Tensor, Tensor]:
    """Per-output-index source taps (4) and Keys-cubic (A=-0.75) weights, following ATen.

    float32: scale and source index are computed in fp32 with the multiply-add of
    ``scale*(dst+0.5)-0.5`` fused (what both ATen's CPU and CUDA builds do).
    """
    i = torch.arange(out_size, dtype=torch.float64)
    if dtype == torch.float32:
        scale = (torch.tensor(float(in_size), dtype=torch.float32) / out_size)
        ip = (torch.arange(out_size, dtype=torch.float32) + 0.5)
        src = (scale.double() * ip.double() - 0.5).float()       # fp32 FMA
    else:
        scale = torch.tensor(float(in_size), dtype=dtype) / out_size
        src = scale * (i.to(dtype) + 0.5) - 0.5
    fl = torch.floor(src)
    idx = fl.long().clamp(max=in_size - 1)
    t = (src - idx.to(src.dtype)).clamp(0, 1)
    A = -0.75

    def c1(x):
        return ((A + 2) * x - (A + 3)) * x * x + 1

    def c2(x):
        return ((A * x - 5 * A) * x + 8 * A) * x - 4 * A

    w = torch.stack([c2(t + 1.0), c1(t), c1(1.0 - t), c2((1.0 - t) + 1.0)], 0).to(dtype)
    taps = torch.stack([(idx + j - 1).clamp(0, in_size - 1) for j in range(4)], 0)
    return taps, w


def bicubic_nchw(x: Tensor, size: Tuple[int, int]) -> Tensor:
    """F.interpolate(x, size, mode='bicubic', align_corners=False) restated (NCHW)."""
    B, C, H, W = x.shape
    oh, ow = int(size[0]), int(size[1])
    ty, wy = _cubic_taps(H, oh, x.dtype)
    tx, wx = _cubic_taps(W, ow, x.dtype)
    out = torch.zeros(B, C, oh, ow, dtype=x.dtype)
    for i in range(4):
        rows = x[:, :, ty[i], :]
        acc = torch.zeros(B, C, oh, ow, dtype=x.dtype)
        for j in range(4):
            acc += rows[:, :, :, tx[j]] * wx[j][None, None, None, :]
        out += acc * wy[i][None, None, :, None]
    return out


def layer_norm(x: Tensor, w: Tensor, b: Tensor, eps: float = 1e-5) -> Tensor:
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + b


def gelu_erf(x: Tensor) -> Tensor:
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def linear(x: Tensor, w: Tensor, b: Optional[Tensor]) -> Tensor:
    y = x.matmul(w.t())
    return y if b is None else y + b


def dense_rel_bias(table: Tensor, ws: int = 8) -> Tensor:
    """(heads, N, N) bias with bias[h,i,j] = table[(yi-yj+ws-1)*(2ws-1) + (xi-xj+ws-1), h]."""
    ys, xs = torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")
    ys, xs = ys.flatten(), xs.flatten()
    idx = (ys[:, None] - ys[None, :] + ws - 1) * (2 * ws - 1) + (xs[:, None] - xs[None, :] + ws - 1)
    return table[idx.reshape(-1)].reshape(ws * ws, ws * ws, -1).permute(2, 0, 1).contiguous()


def window_block(x: Tensor, sd: SD, pre: str, heads: int, dtype) -> Tensor:
    """One WindowTransformerBlock on (nWin, 64, dim) tokens (eval mode: dropout = identity)."""
    nW, N, D = x.shape
    hd = D // heads
    h = layer_norm(x, _p(sd, pre + "norm1.weight", dtype), _p(sd, pre + "norm1.bias", dtype))
    qkv = linear(h, _p(sd, pre + "attn.qkv.weight", dtype), _p(sd, pre + "attn.qkv.bias", dtype))
    qkv = qkv.reshape(nW, N, 3, heads, hd).permute(2, 0, 3, 1, 4)
    q, k, v = qkv[0] * (hd ** -0.5), qkv[1], qkv[2]
    attn = q.matmul(k.transpose(-2, -1)) + dense_rel_bias(_p(sd, pre + "attn.relative_position_bias_table", dtype))[None]
    attn = torch.softmax(attn, dim=-1)
    o = attn.matmul(v).transpose(1, 2).reshape(nW, N, D)
    x = x + linear(o, _p(sd, pre + "attn.proj.weight", dtype), _p(sd, pre + "attn.proj.bias", dtype))
    h = layer_norm(x, _p(sd, pre + "norm2.weight", dtype), _p(sd, pre + "norm2.bias", dtype))
    h = gelu_erf(linear(h, _p(sd, pre + "mlp.0.weight", dtype), _p(sd, pre + "mlp.0.bias", dtype)))
    return x + linear(h, _p(sd, pre + "mlp.2.weight", dtype), _p(sd, pre + "mlp.2.bias", dtype))


def patch_embed_nhwc(feat: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """Conv2d(k=8,s=8) as a GEMM: tokens (B,Ht,Wt,dim) from NHWC feat (floor division of H, W)."""
    B, H, W, C = feat.shape
    Ht, Wt = H // 8, W // 8
    a = feat[:, :Ht * 8, :Wt * 8].reshape(B, Ht, 8, Wt, 8, C).permute(0, 1, 3, 2, 4, 5).reshape(B * Ht * Wt, 64 * C)
    wk = w.permute(0, 2, 3, 1).reshape(w.shape[0], -1)         # (dim, ky*kx*c)
    return (a.matmul(wk.t()) + b).reshape(B, Ht, Wt, -1)


def patch_unembed_nhwc(tok: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """ConvTranspose2d(k=8,s=8) as GEMM + scatter: (B,Ht,Wt,dim) -> NHWC (B,8Ht,8Wt,Cout).

    w is the ConvTranspose2d weight (dim, Cout, 8, 8); out[b, 8ty+ky, 8tx+kx, c] = sum_d tok[b,ty,tx,d] w[d,c,ky,kx] + b[c].
    """
    B, Ht, Wt, D = tok.shape
    Cout = w.shape[1]
    wk = w.permute(0, 2, 3, 1).reshape(D, 64 * Cout)           # (dim, ky*kx*c)
    y = tok.reshape(-1, D).matmul(wk).reshape(B, Ht, Wt, 8, 8, Cout) + b
    return y.permute(0, 1, 3, 2, 4, 5).reshape(B, Ht * 8, Wt * 8, Cout).contiguous()


def window_stack(tokens: Tensor, sd: SD, n_blocks: int, heads: int, dtype, prefix: str = "window_blocks.") -> Tensor:
    """zero-pad token grid to x8, window_partition, blocks, window_reverse, unpad. tokens (B,Ht,Wt,D)."""
    B, Ht, Wt, D = tokens.shape
    Hp, Wp = (Ht + 7) // 8 * 8, (Wt + 7) // 8 * 8
    t = torch.zeros(B, Hp, Wp, D, dtype=dtype)
    t[:, :Ht, :Wt] = tokens                                    # zero pad AFTER the embed bias; no mask
    win = t.reshape(B, Hp // 8, 8, Wp // 8, 8, D).permute(0, 1, 3, 2, 4, 5).reshape(-1, 64, D)
    for i in range(n_blocks):
        win = window_block(win, sd, f"{prefix}{i}.", heads, dtype)
    t = win.reshape(B, Hp // 8, Wp // 8, 8, 8, D).permute(0, 1, 3, 2, 4, 5).reshape(B, Hp, Wp, D)
    return t[:, :Ht, :Wt].contiguous()


def _n_blocks(sd: SD, prefix: str) -> int:
    return 1 + max(int(k[len(prefix):].split(".")[0]) for k in sd if k.startswith(prefix))


# --------------------------------------------------------------------------- WindowTransformer
def window_forward(sd: SD, x: Tensor, res_out=(1080, 1920), upscale_factor=None, dtype=torch.float32,
                   pre_clamp: bool = False) -> Tensor:
    x = x.detach().to("cpu", dtype)
    if upscale_factor is not None:
        res_out = (x.shape[2] * upscale_factor, x.shape[3] * upscale_factor)
    up_in = bicubic_nchw(x, res_out)
    xh = x.permute(0, 2, 3, 1).contiguous()
    feat = conv3x3_nhwc(xh, _p(sd, "conv1.weight", dtype), _p(sd, "conv1.bias", dtype), relu=True)
    feat = conv3x3_nhwc(feat, _p(sd, "conv2.weight", dtype), _p(sd, "conv2.bias", dtype), relu=True)
    fd = conv3x3_nhwc(feat, _p(sd, "downsample.weight", dtype), _p(sd, "downsample.bias", dtype), stride=2)
    tok = patch_embed_nhwc(fd, _p(sd, "patch_embed.weight", dtype), _p(sd, "patch_embed.bias", dtype))
    heads = sd["window_blocks.0.attn.relative_position_bias_table"].shape[1]
    tok = window_stack(tok, sd, _n_blocks(sd, "window_blocks."), heads, dtype)
    ft = patch_unembed_nhwc(tok, _p(sd, "patch_unembed.weight", dtype), _p(sd, "patch_unembed.bias", dtype))
    mh, mw = min(fd.shape[1], ft.shape[1]), min(fd.shape[2], ft.shape[2])
    comb = fd[:, :mh, :mw] + ft[:, :mh, :mw]
    dec = conv3x3_nhwc(comb, _p(sd, "decoder_conv1.weight", dtype), _p(sd, "decoder_conv1.bias", dtype), relu=True)
    res = conv3x3_nhwc(dec, _p(sd, "decoder_conv2.weight", dtype), _p(sd, "decoder_conv2.bias", dtype))
    out = up_in + bicubic_nchw(res.permute(0, 3, 1, 2).contiguous(), res_out)
    return out if pre_clamp else out.clamp(0.0, 1.0)


# --------------------------------------------------------------------------- ResidualTransformer
def mha_block(x: Tensor, sd: SD, pre: str, heads: int, dtype) -> Tensor:
    """ResidualTransformer TransformerBlock: pre-LN, nn.MultiheadAttention(x,x,x), MLP."""
    B, S, D = x.shape
    hd = D // heads
    h = layer_norm(x, _p(sd, pre + "norm1.weight", dtype), _p(sd, pre + "norm1.bias", dtype))
    qkv = linear(h, _p(sd, pre + "attn.in_proj_weight", dtype), _p(sd, pre + "attn.in_proj_bias", dtype))
    q, k, v = [t.reshape(B, S, heads, hd).transpose(1, 2) for t in qkv.split(D, dim=-1)]
    a = torch.softmax(q.matmul(k.transpose(-2, -1)) / math.sqrt(hd), dim=-1)
    o = a.matmul(v).transpose(1, 2).reshape(B, S, D)
    x = x + linear(o, _p(sd, pre + "attn.out_proj.weight", dtype), _p(sd, pre + "attn.out_proj.bias", dtype))
    h = layer_norm(x, _p(sd, pre + "norm2.weight", dtype), _p(sd, pre + "norm2.bias", dtype))
    h = gelu_erf(linear(h, _p(sd, pre + "mlp.0.weight", dtype), _p(sd, pre + "mlp.0.bias", dtype)))
    return x + linear(h, _p(sd, pre + "mlp.2.weight", dtype), _p(sd, pre + "mlp.2.bias", dtype))


def residual_forward(sd: SD, x: Tensor, res_out=(1080, 1920), upscale_factor=None, dtype=torch.float32,
                     heads: int = 8, pre_clamp: bool = False) -> Tensor:
    x = x.detach().to("cpu", dtype)
    if upscale_factor is not None:
        res_out = (x.shape[2] * upscale_factor, x.shape[3] * upscale_factor)
    up_in = bicubic_nchw(x, res_out)
    xh = x.permute(0, 2, 3, 1).contiguous()
    feat = conv3x3_nhwc(xh, _p(sd, "conv1.weight", dtype), _p(sd, "conv1.bias", dtype), relu=True)
    feat = conv3x3_nhwc(feat, _p(sd, "conv2.weight", dtype), _p(sd, "conv2.bias", dtype), relu=True)
    fd = conv3x3_nhwc(feat, _p(sd, "downsample.weight", dtype), _p(sd, "downsample.bias", dtype), stride=2)
    tok = patch_embed_nhwc(fd, _p(sd, "patch_embed.weight", dtype), _p(sd, "patch_embed.bias", dtype))
    B, Ht, Wt, D = tok.shape
    pos = _p(sd, "pos_embed", dtype)
    if Ht * Wt != pos.shape[1]:
        # the reference raises here (ResidualTransformer/model.py:140, broadcast of 3600 fixed tokens)
        raise RuntimeError(f"The size of tensor a ({Ht * Wt}) must match the size of tensor b ({pos.shape[1]}) "
                           "at non-singleton dimension 1")
    t = tok.reshape(B, Ht * Wt, D) + pos
    for i in range(_n_blocks(sd, "transformer_blocks.")):
        t = mha_block(t, sd, f"transformer_blocks.{i}.", heads, dtype)
    ft = patch_unembed_nhwc(t.reshape(B, Ht, Wt, D), _p(sd, "patch_unembed.weight", dtype), _p(sd, "patch_unembed.bias", dtype))
    comb = fd + ft                                           # shapes must agree, as in the reference (:153)
    dec = conv3x3_nhwc(comb, _p(sd, "decoder_conv1.weight", dtype), _p(sd, "decoder_conv1.bias", dtype), relu=True)
    res = conv3x3_nhwc(dec, _p(sd, "decoder_conv2.weight", dtype), _p(sd, "decoder_conv2.bias", dtype))
    out = up_in + bicubic_nchw(res.permute(0, 3, 1, 2).contiguous(), res_out)
    return out if pre_clamp else out.clamp(0.0, 1.0)


# --------------------------------------------------------------------------- FastTransformer
def _upsampler(x: Tensor, sd: SD, pre: str, scale: int, dtype) -> Tensor:
    """FastTransformer/utils.py Upsampler: conv(n->r^2 n)+PixelShuffle(r); scale 4 = two x2 stages."""
    if scale not in (2, 3, 4, 6):
        raise ValueError(f"Requested scale={scale} was not built.")
    stages = [(0, 2), (2, 2)] if scale == 4 else [(0, scale)]
    for idx, r in stages:
        x = conv3x3_nhwc(x, _p(sd, f"{pre}upsamplers.{scale}.{idx}.weight", dtype),
                         _p(sd, f"{pre}upsamplers.{scale}.{idx}.bias", dtype))
        x = pixel_shuffle_nhwc(x, r)
    return x


def reflect_pad_nhwc(x: Tensor, pad_h: int, pad_w: int) -> Tensor:
    """F.pad(mode='reflect') on the bottom/right only: index H+i -> H-2-i."""
    B, H, W, C = x.shape
    iy = torch.cat([torch.arange(H), H - 2 - torch.arange(pad_h)])
    ix = torch.cat([torch.arange(W), W - 2 - torch.arange(pad_w)])
    return x[:, iy][:, :, ix].contiguous()


def aa_bilinear_resize_nchw(x: Tensor, size: Tuple[int, int]) -> Tensor:
    """torchvision Resize(size) on a tensor = F.interpolate(bilinear, antialias=True): separable
    triangle filter with support max(scale,1), weights normalised per output index
    (ATen _upsample_bilinear2d_aa; UpSampleKernel.cpp HelperInterpLinear::aa_filter)."""
    def taps(insz, outsz, dt):
        scale = insz / outsz
        support = scale if scale >= 1.0 else 1.0
        M = torch.zeros(outsz, insz, dtype=dt)
        for i in range(outsz):
            center = scale * (i + 0.5)
            lo = max(int(center - support + 0.5), 0)
            hi = min(int(center + support + 0.5), insz)
            js = torch.arange(lo, hi, dtype=torch.float64)
            inv = 1.0 / scale if scale >= 1.0 else 1.0
            w = (1.0 - ((js - center + 0.5) * inv).abs()).clamp(min=0)
            M[i, lo:hi] = (w / w.sum()).to(dt)
        return M
    My, Mx = taps(x.shape[2], size[0], x.dtype), taps(x.shape[3], size[1], x.dtype)
    return torch.einsum("oh,bchw,pw->bcop", My, x, Mx)


def fast_forward(sd: SD, x: Tensor, res_out=(1080, 1920), upscale_factor=None, require_ratio: bool = True,
                 dtype=torch.float32, pre_clamp: bool = False) -> Tensor:
    x = x.detach().to("cpu", dtype)
    if upscale_factor is not None:
        res_out = (x.shape[2] * upscale_factor, x.shape[3] * upscale_factor)
    else:
        upscale_factor = math.ceil(max(res_out[0] / x.shape[2], res_out[1] / x.shape[3]))
    xh = x.permute(0, 2, 3, 1).contiguous()
    feat = conv3x3_nhwc(xh, _p(sd, "conv1.weight", dtype), _p(sd, "conv1.bias", dtype), relu=True)
    feat = conv3x3_nhwc(feat, _p(sd, "conv2.weight", dtype), _p(sd, "conv2.bias", dtype), relu=True)
    B, H, W, C = feat.shape
    pad_h, pad_w = (8 - H % 8) % 8, (8 - W % 8) % 8
    feat_pad = reflect_pad_nhwc(feat, pad_h, pad_w) if (pad_h or pad_w) else feat
    # branch A: sub-pixel upsample of the features, then 64->3 conv (no bias) + ReLU
    up = _upsampler(feat, sd, "up1.", upscale_factor, dtype)
    up = conv3x3_nhwc(up, _p(sd, "up1_conv.conv.weight", dtype), None, relu=True)
    tok = patch_embed_nhwc(feat_pad, _p(sd, "patch_embed.weight", dtype), _p(sd, "patch_embed.bias", dtype))
    heads = sd["window_blocks.0.attn.relative_position_bias_table"].shape[1]
    tok = window_stack(tok, sd, _n_blocks(sd, "window_blocks."), heads, dtype)
    ft = patch_unembed_nhwc(tok, _p(sd, "patch_unembed.weight", dtype), _p(sd, "patch_unembed.bias", dtype))
    comb = feat + ft[:, :H, :W]
    dec = conv3x3_nhwc(comb, _p(sd, "decoder_conv1.weight", dtype), _p(sd, "decoder_conv1.bias", dtype), relu=True)
    res = conv3x3_nhwc(dec, _p(sd, "decoder_conv2.weight", dtype), _p(sd, "decoder_conv2.bias", dtype))
    # branch B: sub-pixel upsample of the 3-channel residual, then a 3->3 conv
    rup = _upsampler(res, sd, "final_upscale.", upscale_factor, dtype)
    rup = conv3x3_nhwc(rup, _p(sd, "final_upscale_conv.weight", dtype), _p(sd, "final_upscale_conv.bias", dtype))
    out = (up + rup).permute(0, 3, 1, 2).contiguous()
    # the reference compares res_out with (H_out, H_out) (sic, model.py:323) -> Resize is requested for every
    # non-square output; torchvision's Resize returns its input unchanged when the size already matches.
    if require_ratio and tuple(res_out) != (out.shape[2], out.shape[2]):
        if (out.shape[2], out.shape[3]) != (int(res_out[0]), int(res_out[1])):
            out = aa_bilinear_resize_nchw(out, (int(res_out[0]), int(res_out[1])))
    return out if pre_clamp else out.clamp(0.0, 1.0)


FORWARDS = {"WindowTransformer": window_forward, "FastTransformer": fast_forward,
            "ResidualTransformer": residual_forward}


def forward(model: str, sd: SD, x: Tensor, **kw) -> Tensor:
    return FORWARDS[model](sd, x, **kw)
