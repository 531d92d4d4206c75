"""Forward dispatch into libtu_b200 (registered as the PyTorch custom op ``tu::forward``) and thin
single-op wrappers used by the per-op parity tests.

PyTorch is plumbing here: it owns device memory (inputs, outputs, workspace from the caching
allocator) and the CUDA stream; all arithmetic happens in the library's kernels.  CPU tensors are
rejected — there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import itertools
import math
import weakref
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from .packing import PackedWeights

_DT = {torch.float32: _lib.TU_F32, torch.bfloat16: _lib.TU_BF16, torch.uint8: _lib.TU_U8}
_TORCH_DT = {v: k for k, v in _DT.items()}
# handle -> PackedWeights, WEAK: the owner (EngineModel._tu_cache, or a test) keeps the object alive; a handle whose owner is gone
# simply disappears, so a copied / unpickled module can never release another module's weights and nothing leaks
_REGISTRY: "weakref.WeakValueDictionary[int, PackedWeights]" = weakref.WeakValueDictionary()
_next_handle = itertools.count(1)


def register_weights(pw: PackedWeights) -> int:
    h = next(_next_handle)
    _REGISTRY[h] = pw
    return h


def release_weights(handle: int) -> None:
    _REGISTRY.pop(handle, None)


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _require_cuda(*ts: torch.Tensor) -> None:
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("transformerupscaler_b200 runs on CUDA tensors only (no CPU fallback); got a "
                               f"{t.device} tensor")


def _code(dt: torch.dtype) -> int:
    if dt not in _DT:
        raise TypeError(f"unsupported dtype {dt}; the engine handles float32, bfloat16 and (frames only) uint8")
    return _DT[dt]


LAYOUTS = {"chw": 0, "hwc": _lib.TU_LAYOUT_HWC, "hwc_bgr": _lib.TU_LAYOUT_HWC_BGR}


def layout_code(name: str) -> int:
    if name not in LAYOUTS:
        raise ValueError(f"unknown frame layout {name!r}: use 'chw', 'hwc' or 'hwc_bgr'")
    return LAYOUTS[name]


def _forward_impl(x: torch.Tensor, handle: int, out_h: int, out_w: int, scale: int, compute_bf16: bool,
                  out_code: int, clamp: bool, in_layout: int = 0) -> torch.Tensor:
    lib = _lib.load()
    pw = _REGISTRY.get(handle)
    if pw is None:
        raise RuntimeError(f"tu::forward: packed-weights handle {handle} is not alive (its TransformerModel was deleted or repacked)")
    _require_cuda(x)
    if in_layout:
        if x.dim() != 4 or x.shape[3] != 3 or x.dtype != torch.uint8:
            raise RuntimeError(f"interleaved frames must be uint8 of shape (B,H,W,3), got {x.dtype} {tuple(x.shape)}")
        x = x.contiguous()
        B, H, W, _ = x.shape
    else:
        if x.dim() != 4 or x.shape[1] != 3:
            raise RuntimeError(f"expected input of shape (B,3,H,W), got {tuple(x.shape)}")
        x = x.contiguous()
        B, _, H, W = x.shape
    cdt = _lib.TU_BF16 if compute_bf16 else _lib.TU_F32
    out_shape = (B, out_h, out_w, 3) if out_code >> 8 else (B, 3, out_h, out_w)
    if out_code >> 8 and (out_code & 0xFF) != _lib.TU_U8:
        raise RuntimeError("interleaved output layouts are for uint8 frames")
    out = torch.empty(out_shape, dtype=_TORCH_DT[out_code & 0xFF], device=x.device)
    nbytes = lib.tu_forward_workspace_bytes_for(C.byref(pw.struct), B, H, W, out_h, out_w, scale, cdt)
    if nbytes == 0:
        _lib.check(_lib.TU_ERR_SCALE if "was not built" in _lib.last_error() else _lib.TU_ERR_ARG)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.tu_forward(C.byref(pw.struct), x.data_ptr(), _code(x.dtype) | in_layout, out.data_ptr(), out_code,
                            B, H, W, out_h, out_w, scale, cdt, int(clamp), ws.data_ptr(), nbytes, _stream())
    _lib.check(rc)
    return out


@torch.library.custom_op("tu::forward", mutates_args=(), device_types="cuda")
def tu_forward(x: torch.Tensor, handle: int, out_h: int, out_w: int, scale: int, compute_bf16: bool,
               out_code: int, clamp: bool, in_layout: int = 0) -> torch.Tensor:
    return _forward_impl(x, handle, out_h, out_w, scale, compute_bf16, out_code, clamp, in_layout)


@tu_forward.register_fake
def _(x, handle, out_h, out_w, scale, compute_bf16, out_code, clamp, in_layout=0):
    shape = (x.shape[0], out_h, out_w, 3) if out_code >> 8 else (x.shape[0], 3, out_h, out_w)
    return x.new_empty(shape, dtype=_TORCH_DT[out_code & 0xFF])


@torch.library.custom_op("tu::resize_aa", mutates_args=(), device_types="cuda")
def tu_resize_aa(x: torch.Tensor, out_h: int, out_w: int, clamp: bool, out_code: int = -1) -> torch.Tensor:
    """antialiased bilinear Resize of a (B,3,H,W) float image; out_code < 0: same dtype as x, else a dtype (| layout) code"""
    lib = _lib.load()
    _require_cuda(x)
    x = x.contiguous()
    B, Cc, H, W = x.shape
    if out_code < 0:
        out_code = _code(x.dtype)
    if Cc != 3 and out_code >> 8:
        raise RuntimeError("interleaved output needs a 3-channel image")
    shape = (B, out_h, out_w, 3) if out_code >> 8 else (B, Cc, out_h, out_w)
    out = torch.empty(shape, dtype=_TORCH_DT[out_code & 0xFF], device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.tu_resize_bilinear_aa_to(x.data_ptr(), _code(x.dtype), out.data_ptr(), out_code, B * Cc // 3, H, W, out_h, out_w,
                                                int(clamp), _stream()))
    return out


@tu_resize_aa.register_fake
def _(x, out_h, out_w, clamp, out_code=-1):
    if out_code < 0:
        return x.new_empty((x.shape[0], x.shape[1], out_h, out_w))
    shape = (x.shape[0], out_h, out_w, 3) if out_code >> 8 else (x.shape[0], x.shape[1], out_h, out_w)
    return x.new_empty(shape, dtype=_TORCH_DT[out_code & 0xFF])


def run_forward(pw_handle: int, model: str, x: torch.Tensor, res_out: Tuple[int, int], upscale_factor: Optional[int],
                require_ratio: bool, compute_bf16: bool, out_dtype: torch.dtype, clamp: bool = True,
                in_layout: str = "chw", out_layout: str = "chw") -> torch.Tensor:
    """Shape logic of the three reference forwards (W:237-238, F:245-248,323-327, R:121-122) around tu::forward.
    in_layout / out_layout ('chw' | 'hwc' | 'hwc_bgr'): uint8 frames may be interleaved (B,H,W,3) on either side."""
    il, ol = layout_code(in_layout), layout_code(out_layout)
    H, W = (int(x.shape[1]), int(x.shape[2])) if il else (int(x.shape[2]), int(x.shape[3]))
    if upscale_factor is not None:
        res_out = (H * upscale_factor, W * upscale_factor)
    res_out = (int(res_out[0]), int(res_out[1]))
    out_code = _code(out_dtype) | ol
    if model != "FastTransformer":
        return tu_forward(x, pw_handle, res_out[0], res_out[1], 0, compute_bf16, out_code, clamp, il)
    scale = upscale_factor if upscale_factor is not None else math.ceil(max(res_out[0] / H, res_out[1] / W))
    if scale not in (2, 3, 4, 6):
        raise ValueError(f"Requested scale={scale} was not built.")
    oh, ow = H * scale, W * scale
    # the reference tests res_out against (H_out, H_out) (sic): Resize is requested for every non-square
    # output and is the identity when the size already matches
    need_resize = require_ratio and res_out != (oh, oh) and res_out != (oh, ow)
    if not need_resize:
        return tu_forward(x, pw_handle, oh, ow, scale, compute_bf16, out_code, clamp, il)
    # Resize is the last op: the un-clamped full-size image stays in the compute dtype and the Resize kernel writes the requested
    # output dtype / layout (uint8 frames: (out*255).clamp(0,255).to(uint8) fused into its store)
    mid_code = _code(torch.bfloat16 if compute_bf16 and out_dtype != torch.float32 else torch.float32)
    full = tu_forward(x, pw_handle, oh, ow, scale, compute_bf16, mid_code, False, il)
    return tu_resize_aa(full, res_out[0], res_out[1], clamp, out_code)
