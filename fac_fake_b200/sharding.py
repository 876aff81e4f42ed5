"""Whole-video sharding across the GPUs of one box.

A video's score depends only on its own crops, so rank r owns a contiguous block of whole
videos, weights are replicated and NO collective sits on the data path; the only exchange is
the final gather of the per-video fp32 scores (<= 32 KB for 8192 videos).
"""
from __future__ import annotations

from typing import List, Sequence, Tuple

import torch


def shard_range(n_videos: int, rank: int, world: int) -> Tuple[int, int]:
    """Equal-count contiguous block [lo, hi) of rank `rank`."""
    base, rem = divmod(n_videos, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_by_crops(crop_counts: Sequence[int], world: int) -> List[Tuple[int, int]]:
    """Contiguous blocks of whole videos balanced by crop count (videos have unequal face counts)."""
    total = sum(crop_counts)
    bounds, acc, v = [], 0, 0
    n = len(crop_counts)
    for r in range(world):
        lo = v
        target = total * (r + 1) / world
        while v < n and (acc + crop_counts[v] / 2.0 <= target or r == world - 1):
            acc += crop_counts[v]
            v += 1
        # leave at least enough videos for the remaining ranks only if any remain
        bounds.append((lo, v))
    bounds[-1] = (bounds[-1][0], n)
    return bounds


def gather_scores(local_scores: torch.Tensor, n_videos: int, rank: int, world: int) -> torch.Tensor:
    """All ranks receive the full [n_videos] score vector (torch.distributed; nccl or gloo)."""
    if world == 1:
        return local_scores
    import torch.distributed as dist
    base, rem = divmod(n_videos, world)
    width = base + (1 if rem else 0)
    pad = torch.zeros((width,), dtype=torch.float32, device=local_scores.device)
    pad[: local_scores.numel()] = local_scores
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad)
    parts = []
    for r in range(world):
        lo, hi = shard_range(n_videos, r, world)
        parts.append(bufs[r][: hi - lo])
    return torch.cat(parts)
