"""Forward dispatch into libtu_b200 (registered as the PyTorch custom op ``tu::forward``) and thin
single-op wrappers used by the per-op parity tests.

PyTorch is plumbing here: it owns device memory (inputs, outputs, workspace from the caching
allocator) and the CUDA stream; all arithmetic happens in the library's kernels.  CPU tensors are
rejected — there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from .packing import PackedWeights

_DT = {torch.float32: _lib.TU_F32, torch.bfloat16: _lib.TU_BF16, torch.uint8: _lib.TU_U8}
_TORCH_DT = {v: k for k, v in _DT.items()}
_REGISTRY: Dict[int, PackedWeights] = {}
_next_handle = [1]


def register_weights(pw: PackedWeights) -> int:
    h = _next_handle[0]
    _next_handle[0] += 1
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


def _forward_impl(x: torch.Tensor, handle: int, out_h: int, out_w: int, scale: int, compute_bf16: bool,
                  out_code: int, clamp: bool) -> torch.Tensor:
    lib = _lib.load()
    pw = _REGISTRY[handle]
    _require_cuda(x)
    if x.dim() != 4 or x.shape[1] != 3:
        raise RuntimeError(f"expected input of shape (B,3,H,W), got {tuple(x.shape)}")
    x = x.contiguous()
    B, _, H, W = x.shape
    cdt = _lib.TU_BF16 if compute_bf16 else _lib.TU_F32
    out = torch.empty((B, 3, out_h, out_w), dtype=_TORCH_DT[out_code], device=x.device)
    mid = _lib.MODEL_IDS[pw.model]
    nbytes = lib.tu_forward_workspace_bytes(mid, B, H, W, out_h, out_w, scale, cdt)
    if nbytes == 0:
        _lib.check(_lib.TU_ERR_SCALE if "was not built" in _lib.last_error() else _lib.TU_ERR_ARG)
    ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = lib.tu_forward(C.byref(pw.struct), x.data_ptr(), _code(x.dtype), out.data_ptr(), _code(out.dtype),
                            B, H, W, out_h, out_w, scale, cdt, int(clamp), ws.data_ptr(), nbytes, _stream())
    _lib.check(rc)
    return out


@torch.library.custom_op("tu::forward", mutates_args=(), device_types="cuda")
def tu_forward(x: torch.Tensor, handle: int, out_h: int, out_w: int, scale: int, compute_bf16: bool,
               out_code: int, clamp: bool) -> torch.Tensor:
    return _forward_impl(x, handle, out_h, out_w, scale, compute_bf16, out_code, clamp)


@tu_forward.register_fake
def _(x, handle, out_h, out_w, scale, compute_bf16, out_code, clamp):
    return x.new_empty((x.shape[0], 3, out_h, out_w), dtype=_TORCH_DT[out_code])


@torch.library.custom_op("tu::resize_aa", mutates_args=(), device_types="cuda")
def tu_resize_aa(x: torch.Tensor, out_h: int, out_w: int, clamp: bool) -> torch.Tensor:
    lib = _lib.load()
    _require_cuda(x)
    x = x.contiguous()
    B, Cc, H, W = x.shape
    out = torch.empty((B, Cc, out_h, out_w), dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        _lib.check(lib.tu_resize_bilinear_aa(x.data_ptr(), _code(x.dtype), out.data_ptr(), B * Cc // 3, H, W, out_h, out_w,
                                             int(clamp), _stream()))
    return out


@tu_resize_aa.register_fake
def _(x, out_h, out_w, clamp):
    return x.new_empty((x.shape[0], x.shape[1], out_h, out_w))


def run_forward(pw_handle: int, model: str, x: torch.Tensor, res_out: Tuple[int, int], upscale_factor: Optional[int],
                require_ratio: bool, compute_bf16: bool, out_dtype: torch.dtype, clamp: bool = True) -> torch.Tensor:
    """Shape logic of the three reference forwards (W:237-238, F:245-248,323-327, R:121-122) around tu::forward."""
    H, W = int(x.shape[2]), int(x.shape[3])
    if upscale_factor is not None:
        res_out = (H * upscale_factor, W * upscale_factor)
    res_out = (int(res_out[0]), int(res_out[1]))
    out_code = _code(out_dtype)
    if model != "FastTransformer":
        return tu_forward(x, pw_handle, res_out[0], res_out[1], 0, compute_bf16, out_code, clamp)
    scale = upscale_factor if upscale_factor is not None else math.ceil(max(res_out[0] / H, res_out[1] / W))
    if scale not in (2, 3, 4, 6):
        raise ValueError(f"Requested scale={scale} was not built.")
    oh, ow = H * scale, W * scale
    # the reference tests res_out against (H_out, H_out) (sic): Resize is requested for every non-square
    # output and is the identity when the size already matches
    need_resize = require_ratio and res_out != (oh, oh) and res_out != (oh, ow)
    if not need_resize:
        return tu_forward(x, pw_handle, oh, ow, scale, compute_bf16, out_code, clamp)
    if out_dtype == torch.uint8:
        raise NotImplementedError("uint8 frame output is not available on FastTransformer's antialiased-Resize path "
                                  "(res_out that is not an integer multiple of the input): request float output")
    full = tu_forward(x, pw_handle, oh, ow, scale, compute_bf16, out_code, False)
    return tu_resize_aa(full, res_out[0], res_out[1], clamp)
