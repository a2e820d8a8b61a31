"""Sharding of independent video streams over ranks (one process per GPU) and whole-job timing.

Frames of one clip are sequentially dependent through the Q/K/V FIFO (td4_psp18.py:145-154), clips
are not: the path shards over GPUs by giving every rank its own clip(s) and its own model replica.
No tensor crosses NVLink; torch.distributed is used only for the launch barriers and to combine
per-rank (frames, elapsed) into whole-job throughput = total frames / max elapsed.
Backend 'nccl' on the GPU box, 'gloo' in the CPU tests.
"""
from __future__ import annotations

from typing import List


def clips_for_rank(rank: int, world: int, n_clips: int) -> List[int]:
    """Round-robin ownership: rank r runs clips {r, r + world, ...}."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    return list(range(rank, n_clips, world))


def whole_job_throughput(frames_local: int, elapsed_ms_local: float, group=None, device=None):
    """(total frames over all ranks, max elapsed ms over ranks, frames/s).  Collective: every rank of
    `group` must call it.  Falls back to the local numbers when torch.distributed is not initialised."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()):
        return frames_local, elapsed_ms_local, frames_local / (elapsed_ms_local / 1e3)
    t = torch.tensor([float(frames_local), 0.0], dtype=torch.float64, device=device)
    m = torch.tensor([float(elapsed_ms_local)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    dist.all_reduce(m, op=dist.ReduceOp.MAX, group=group)
    total, worst = float(t[0]), float(m[0])
    return int(round(total)), worst, total / (worst / 1e3)


def barrier(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        dist.barrier(group=group)


def partition_is_exact(world: int, n_clips: int) -> bool:
    """Every clip is owned by exactly one rank."""
    seen: List[int] = []
    for r in range(world):
        seen += clips_for_rank(r, world, n_clips)
    return sorted(seen) == list(range(n_clips))
