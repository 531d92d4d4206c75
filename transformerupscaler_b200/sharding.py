"""Frame sharding across GPUs.  The workload is a batch of independent video frames (no cross-frame state in any of
the three models), so the path shards by frame with NO data-path collective; torch.distributed only carries the
barrier, the max-over-ranks of the device timings and — outside the timed region — an optional gather of outputs
(rank r of N takes frames [r*B/N, (r+1)*B/N))."""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def frame_shard(total_frames: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, stop) slice of frames for `rank`; the first (total % world) ranks take one extra frame."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(total_frames, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def max_over_ranks(value: float, device=None) -> float:
    """max over ranks of a per-rank scalar (e.g. CUDA-event milliseconds); identity when not distributed."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_frames(local: torch.Tensor, total_frames: int) -> torch.Tensor:
    """All-gather the per-rank output frames back into batch order (ragged shards are padded to the largest)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [frame_shard(total_frames, r, world) for r in range(world)]
    biggest = max(b - a for a, b in sizes)
    pad = torch.zeros((biggest,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[: local.shape[0]] = local
    parts: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad)
    return torch.cat([p[: b - a] for p, (a, b) in zip(parts, sizes)], dim=0)
