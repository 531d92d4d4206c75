"""Helpers for the GPU parity tests: call single ops of libtu_b200 through its C ABI on torch CUDA tensors."""
import ctypes as C

import torch

from transformerupscaler_b200 import _lib

DT = {torch.float32: 0, torch.bfloat16: 1, torch.uint8: 2}


def stream():
    return torch.cuda.current_stream().cuda_stream


def p(t):
    return 0 if t is None else t.data_ptr()


def chk(rc):
    _lib.check(rc)


def stem_conv(x, w27x64, b, dtype, w64=None):
    lib = _lib.load()
    B, _, H, W = x.shape
    out = torch.empty(B, H, W, 64, dtype=dtype, device=x.device)
    chk(lib.tu_stem_conv(p(x), DT[x.dtype], p(w27x64), p(w64), p(b), p(out), DT[dtype], B, H, W, stream()))
    return out


def conv3x3_c64(x, w, b, stride=1, relu=0, nchunk=1, ps_r=0):
    lib = _lib.load()
    B, H, W, _ = x.shape
    Ho, Wo = (H - 1) // stride + 1, (W - 1) // stride + 1
    if ps_r:
        out = torch.empty(B, Ho * ps_r, Wo * ps_r, 64, dtype=x.dtype, device=x.device)
    else:
        out = torch.empty(B, Ho, Wo, 64 * nchunk, dtype=x.dtype, device=x.device)
    chk(lib.tu_conv3x3_c64(p(x), p(w), p(b), p(out), DT[x.dtype], B, H, W, stride, relu, nchunk, ps_r, stream()))
    return out


def conv64to3(x, w, b, relu=0, w16=None):
    lib = _lib.load()
    B, H, W, _ = x.shape
    out = torch.empty(B, 3, H, W, dtype=torch.float32, device=x.device)
    chk(lib.tu_conv3x3_c64_to3(p(x), DT[x.dtype], p(w), p(w16), p(b), p(out), B, H, W, relu, stream()))
    return out


def conv64to3_stream(x, w, b, relu=0):
    """x NHWC bf16; w (3, 64, 3, 3), b (3) or None: packed here the way packing.PackedWeights does"""
    lib = _lib.load()
    B, H, W, _ = x.shape
    t = torch.zeros(3, 3, 16, 64)
    t[:, :, :3] = w.float().permute(3, 2, 0, 1).flip(1)
    b16 = torch.zeros(16)
    if b is not None:
        b16[:3] = b
    wst, b16 = t.to(x.device, torch.bfloat16).contiguous(), b16.to(x.device)
    out = torch.empty(B, 3, H, W, dtype=torch.float32, device=x.device)
    chk(lib.tu_conv3x3_c64_to3_stream(p(x), p(wst), p(b16), p(out), B, H, W, relu, stream()))
    torch.cuda.synchronize()
    return out


def conv3_ps(x, w, b, r):
    lib = _lib.load()
    B, _, H, W = x.shape
    out = torch.empty(B, 3, H * r, W * r, dtype=torch.float32, device=x.device)
    chk(lib.tu_conv3x3_c3_ps(p(x), p(w), p(b), p(out), B, H, W, r, stream()))
    return out


def final_conv_add(x, w, b, addend, out_dtype, clamp):
    lib = _lib.load()
    B, _, H, W = x.shape
    out = torch.empty(B, 3, H, W, dtype=out_dtype, device=x.device)
    chk(lib.tu_final_conv_add(p(x), p(w), p(b), p(addend), p(out), DT[out_dtype], B, H, W, int(clamp), stream()))
    return out


def conv12_fused(x, w64, b1, w2, b2):
    lib = _lib.load()
    B, _, H, W = x.shape
    dt = 2 if x.dtype == torch.uint8 else DT[x.dtype]
    out = torch.empty(B, H, W, 64, dtype=torch.bfloat16, device=x.device)
    chk(lib.tu_conv12_fused(p(x), dt, p(w64), p(b1), p(w2), p(b2), p(out), B, H, W, stream()))
    return out


def dec12_fused(x, w1p, b1, w16, b2):
    """x NHWC bf16; w1p (9, 64, 64) bf16 tap-major; w16 (3, 16, 64) bf16 head bank; planar fp32 (B, 3, H, W) out"""
    lib = _lib.load()
    B, H, W, _ = x.shape
    out = torch.full((B, 3, H, W), float("nan"), dtype=torch.float32, device=x.device)      # the op must write / zero every pixel
    chk(lib.tu_dec12_fused(p(x), p(w1p), p(b1), p(w16), p(b2), p(out), B, H, W, stream()))
    torch.cuda.synchronize()
    return out


def upfold_conv(x, w1, b1, w2, r):
    """x NHWC bf16 on the GPU; folds (w1, b1, w2) on the host like packing.PackedWeights does"""
    from transformerupscaler_b200.packing import fold_up1, pack_fold_bank
    lib = _lib.load()
    B, H, W, _ = x.shape
    Wf, bf = fold_up1(w1, b1, w2, r)
    bank, bias = pack_fold_bank(Wf, bf, r)
    keep = [bank.to(x.device, torch.bfloat16).contiguous(), bias.to(x.device, torch.float32).contiguous(),
            Wf.permute(0, 1, 2, 4, 5, 3).reshape(9, 3 * r * r, 25, 64).to(x.device, torch.float32).contiguous(),
            bf.reshape(9, 3 * r * r).to(x.device, torch.float32).contiguous()]
    f = _lib.TuUpFold()
    f.w, f.b, f.ring_w, f.ring_b, f.r = p(keep[0]), p(keep[1]), p(keep[2]), p(keep[3]), r
    out = torch.empty(B, 3, H * r, W * r, dtype=torch.float32, device=x.device)
    chk(lib.tu_upfold_conv(p(x), C.byref(f), p(out), B, H, W, stream()))
    torch.cuda.synchronize()
    return out, Wf, bf


def subpixel_conv_add(x, w_ps, b_ps, r, fin_wb_host, addend, out_dtype, clamp):
    """fin_wb_host: CPU float tensor of 84 values ((27,3) filter then bias) -- a HOST pointer in the C ABI"""
    lib = _lib.load()
    B, _, H, W = x.shape
    out = torch.empty(B, 3, H * r, W * r, dtype=out_dtype, device=x.device)
    host = (C.c_float * 84)(*fin_wb_host.reshape(-1).tolist())
    chk(lib.tu_subpixel_conv_add(p(x), p(w_ps), p(b_ps), r, C.cast(host, C.c_void_p), p(addend), p(out), DT[out_dtype], B, H, W,
                                 int(clamp), stream()))
    return out


def patch_embed(feat, w, b, pos, Ht, Wt, dim, window, reflect):
    lib = _lib.load()
    B, H, W, _ = feat.shape
    if window:
        M = B * ((Ht + 7) // 8) * ((Wt + 7) // 8) * 64
    else:
        M = B * Ht * Wt
    tok = torch.zeros(M, dim, dtype=torch.float32, device=feat.device)
    chk(lib.tu_patch_embed(p(feat), DT[feat.dtype], p(w), p(b), p(pos), p(tok), B, H, W, Ht, Wt, dim, int(window),
                           int(reflect), stream()))
    return tok


def patch_unembed(tok, w, b, skip, B, Ht, Wt, Hc, Wc, dim, window):
    lib = _lib.load()
    out = torch.empty(B, Hc, Wc, 64, dtype=skip.dtype, device=skip.device)
    chk(lib.tu_patch_unembed(p(tok), p(w), p(b), p(skip), skip.shape[1], skip.shape[2], p(out), DT[skip.dtype], B, Ht, Wt,
                             Hc, Wc, dim, int(window), stream()))
    return out


def transformer_block(x, bw, dim, heads, window, S, dtype, full_workspace=False):
    """full_workspace: size the workspace with tu_block_workspace_bytes_for (enables the tcgen05 global attention)"""
    lib = _lib.load()
    M = x.shape[0]
    n = (lib.tu_block_workspace_bytes_for(M, dim, DT[dtype], int(window), S) if full_workspace
         else lib.tu_block_workspace_bytes(M, dim, DT[dtype]))
    ws = torch.empty(n, dtype=torch.uint8, device=x.device)
    chk(lib.tu_transformer_block(p(x), C.byref(bw), M, dim, heads, int(window), S, DT[dtype], p(ws), n, stream()))
    return x


def bicubic_add_clamp(x, res, outH, outW, out_dtype, clamp, layout=0):
    """layout: 0 planar (B,3,H,W); _lib.TU_LAYOUT_HWC / TU_LAYOUT_HWC_BGR: interleaved uint8 frames (B,H,W,3)"""
    lib = _lib.load()
    B, _, H, W = x.shape
    shape = (B, 3, outH, outW) if layout == 0 else (B, outH, outW, 3)
    out = torch.empty(shape, dtype=out_dtype, device=x.device)
    rH, rW = (res.shape[2], res.shape[3]) if res is not None else (0, 0)
    chk(lib.tu_bicubic_add_clamp(p(x), DT[x.dtype], H, W, p(res), rH, rW, p(out), DT[out_dtype] | layout, B, outH, outW,
                                 int(clamp), stream()))
    return out


def resize_aa(x, outH, outW, clamp):
    lib = _lib.load()
    B, _, H, W = x.shape
    out = torch.empty(B, 3, outH, outW, dtype=x.dtype, device=x.device)
    chk(lib.tu_resize_bilinear_aa(p(x), DT[x.dtype], p(out), B, H, W, outH, outW, int(clamp), stream()))
    return out


def window_attention(qkv, rel_bias, dim, heads):
    lib = _lib.load()
    M = qkv.shape[0]
    out = torch.empty(M, dim, dtype=qkv.dtype, device=qkv.device)
    chk(lib.tu_window_attention(p(qkv), p(rel_bias), p(out), M // 64, dim, heads, DT[qkv.dtype], stream()))
    return out


def window_stack(tok, pw):
    lib = _lib.load()
    chk(lib.tu_window_stack(p(tok), C.byref(pw.struct), tok.shape[0], stream()))
    return tok
