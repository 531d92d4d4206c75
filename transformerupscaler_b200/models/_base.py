"""Shared scaffolding of the drop-in model classes.

The classes below only *hold parameters* under the reference's names and shapes so that
``load_state_dict(strict=True)`` of reference checkpoints works; none of their
``forward`` methods does arithmetic.  ``EngineModel.forward`` routes the whole forward pass into
libtu_b200 through the ``tu::forward`` custom op.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.nn as nn

from .. import engine
from ..packing import PackedWeights


def _relative_position_index(ws: int) -> torch.Tensor:
    ys, xs = torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")
    ys, xs = ys.reshape(-1), xs.reshape(-1)
    return (ys[:, None] - ys[None, :] + ws - 1) * (2 * ws - 1) + (xs[:, None] - xs[None, :] + ws - 1)


class _ParamsOnly(nn.Module):
    def forward(self, *a, **k):  # pragma: no cover - never used
        raise RuntimeError("this sub-module only stores parameters; call TransformerModel.forward")


class WindowAttention(_ParamsOnly):
    """Parameter holder for reference WindowAttention (WindowTransformer/model.py:63-100)."""

    def __init__(self, dim: int, window_size: int, num_heads: int, dropout: float = 0.0):
        super().__init__()
        assert (dim // num_heads) * num_heads == dim, "dim must be divisible by num_heads"
        self.dim, self.window_size, self.num_heads = dim, window_size, num_heads
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.attn_drop = nn.Dropout(dropout)
        self.proj = nn.Linear(dim, dim)
        self.proj_drop = nn.Dropout(dropout)
        self.relative_position_bias_table = nn.Parameter(torch.zeros((2 * window_size - 1) ** 2, num_heads))
        self.register_buffer("relative_position_index", _relative_position_index(window_size))
        nn.init.trunc_normal_(self.relative_position_bias_table, std=0.02)


def _mlp(dim: int, hidden: int, dropout: float) -> nn.Sequential:
    return nn.Sequential(nn.Linear(dim, hidden), nn.GELU(), nn.Linear(hidden, dim), nn.Dropout(dropout))


class WindowTransformerBlock(_ParamsOnly):
    """Parameter holder for reference WindowTransformerBlock (WindowTransformer/model.py:133-149)."""

    def __init__(self, dim: int, window_size: int, num_heads: int, mlp_ratio: float = 4.0, dropout: float = 0.1):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim)
        self.attn = WindowAttention(dim, window_size, num_heads, dropout)
        self.norm2 = nn.LayerNorm(dim)
        self.mlp = _mlp(dim, int(dim * mlp_ratio), dropout)


class GlobalTransformerBlock(_ParamsOnly):
    """Parameter holder for ResidualTransformer's TransformerBlock (ResidualTransformer/model.py:22-38)."""

    def __init__(self, embed_dim: int, num_heads: int, mlp_ratio: float = 4.0, dropout: float = 0.1):
        super().__init__()
        self.norm1 = nn.LayerNorm(embed_dim)
        self.attn = nn.MultiheadAttention(embed_dim, num_heads, dropout=dropout, batch_first=True)
        self.norm2 = nn.LayerNorm(embed_dim)
        self.mlp = _mlp(embed_dim, int(embed_dim * mlp_ratio), dropout)


def _invalidate_after_load(module, incompatible_keys) -> None:
    module.invalidate_engine_cache()


class EngineModel(nn.Module):
    """Base of the three TransformerModel classes: precision selection, weight packing cache, dispatch."""

    ENGINE_MODEL = ""          # "WindowTransformer" | "FastTransformer" | "ResidualTransformer"
    AUTOCAST_OUT_FP32 = True   # reference output dtype under autocast: fp32 (Window/Residual), low precision (Fast)

    def __init__(self):
        super().__init__()
        # (compute_bf16, device) -> (key, PackedWeights, registry handle).  The PackedWeights objects are owned HERE; the engine's
        # registry only holds weak references, so deleting the module (or dropping the entry) frees the packed GPU tensors.
        self._tu_cache = {}
        self._tu_tensors = None      # cached [parameters..., buffers...] (walking the module tree costs ~0.5 ms per forward)
        # None = follow the reference's dtype semantics; "fp32" / "bf16" force the compute path
        self.engine_precision: Optional[str] = None
        self.register_load_state_dict_post_hook(_invalidate_after_load)

    # -- packing ------------------------------------------------------------------------------
    def invalidate_engine_cache(self) -> None:
        """Forget the packed weights (they are rebuilt by the next forward).  Called automatically by load_state_dict and by
        .to() / .cuda() / .float() / .bfloat16(); in-place edits of a parameter are detected through its version counter.  Only
        code that assigns NEW Parameter objects to sub-modules has to call this itself."""
        self._tu_cache = {}
        self._tu_tensors = None

    def _apply(self, fn, *a, **k):
        out = super()._apply(fn, *a, **k)
        self.invalidate_engine_cache()
        return out

    # a copy.deepcopy / pickle of the module must not share the original's packed weights (both go through __getstate__)
    def __getstate__(self):
        st = dict(self.__dict__)
        st["_tu_cache"], st["_tu_tensors"] = {}, None
        return st

    def _packed(self, compute_bf16: bool, device: torch.device) -> int:
        if self._tu_tensors is None:
            self._tu_tensors = list(self.parameters()) + list(self.buffers())
        key = tuple([(t.data_ptr(), t._version) for t in self._tu_tensors])
        slot = (bool(compute_bf16), str(device))
        hit = self._tu_cache.get(slot)
        if hit is None or hit[0] != key:
            sd = {k: v for k, v in self.state_dict(keep_vars=True).items()}
            pw = PackedWeights(self.ENGINE_MODEL, sd, torch.bfloat16 if compute_bf16 else torch.float32, device)
            # The repacking kernels run on the caller's current stream.  Forwards may follow on OTHER streams (FramePipeline alternates
            # compute streams), so the packed tensors are complete before anyone can use them: one host wait per state_dict version.
            if torch.device(device).type == "cuda":
                torch.cuda.current_stream(device).synchronize()
            hit = (key, pw, engine.register_weights(pw))
            self._tu_cache[slot] = hit
        return hit[2]

    def _select_precision(self, x: torch.Tensor) -> Tuple[bool, torch.dtype]:
        """(compute in bf16?, output dtype) mirroring what the reference returns for this input/autocast state."""
        pdt = self.conv1.weight.dtype
        if pdt not in (torch.float32, torch.bfloat16):
            raise TypeError(f"unsupported parameter dtype {pdt}: the engine runs fp32 or bf16 modules")
        autocast = x.is_cuda and torch.is_autocast_enabled("cuda")
        if self.engine_precision == "fp32":
            bf16 = False
        elif self.engine_precision == "bf16":
            bf16 = True
        else:
            bf16 = autocast or pdt == torch.bfloat16
        if pdt == torch.bfloat16:
            out_dt = torch.bfloat16
        elif autocast and not self.AUTOCAST_OUT_FP32:
            out_dt = torch.bfloat16
        else:
            out_dt = torch.float32
        return bf16, out_dt

    def forward(self, x: torch.Tensor, res_out: Tuple[int, int] = (1080, 1920), upscale_factor: int = None,
                require_ratio: bool = True, in_layout: str = "chw", out_layout: str = "chw") -> torch.Tensor:
        """The reference's signature (W:224, F:231, R:114) plus two keyword-only-in-spirit extensions for uint8 video frames:
        in_layout / out_layout in {'chw', 'hwc', 'hwc_bgr'} — uint8 frames may arrive / leave interleaved (B,H,W,3), the form PIL
        and OpenCV hold them in (inference.py:65-70, app_overlay.py:382-386), with the channel swap done in the last kernel."""
        if self.training:
            raise RuntimeError("transformerupscaler_b200 is a forward-only inference engine: call .eval() first "
                               "(dropout is treated as identity; there is no backward)")
        if not x.is_cuda:
            raise RuntimeError("transformerupscaler_b200 has no CPU path: move the model and the input to a CUDA device")
        if x.requires_grad and torch.is_grad_enabled():
            raise RuntimeError("transformerupscaler_b200 is forward-only (no autograd): the input requires grad; call under torch.no_grad()")
        # uint8 frames in -> uint8 frames out (ToTensor scaling on read, (out*255).clamp(0,255).to(uint8) on write, fused into
        # the first and last kernels): the video-pipeline entry, an extension of the reference's float-only signature
        frames_u8 = x.dtype == torch.uint8
        if x.dtype not in (torch.float32, torch.bfloat16, torch.uint8):
            x = x.float()
        bf16, out_dt = self._select_precision(x)
        if frames_u8:
            out_dt = torch.uint8
        handle = self._packed(bf16, x.device)
        if (in_layout != "chw" or out_layout != "chw") and not frames_u8:
            raise ValueError("in_layout / out_layout other than 'chw' apply to uint8 frames only")
        out = engine.run_forward(handle, self.ENGINE_MODEL, x, res_out, upscale_factor, require_ratio, bf16, out_dt,
                                 in_layout=in_layout, out_layout=out_layout)
        if not frames_u8 and x.is_cuda and torch.is_autocast_enabled("cuda") and not self.AUTOCAST_OUT_FP32:
            want = torch.get_autocast_dtype("cuda")
            if out.dtype != want:
                out = out.to(want)
        return out
