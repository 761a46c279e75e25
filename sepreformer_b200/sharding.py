"""Utterance sharding across ranks.

The separator forward has no cross-utterance coupling, so the multi-GPU path is: every rank owns a contiguous block of
utterances, weights are resident per rank, and the only exchange is an all-gather of the small per-utterance result
vector (what ``engine.py:131-146`` accumulates on one device in the reference's ``data_parallel`` flow).
"""
from __future__ import annotations

from typing import List, Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, world: int, rank: int) -> Tuple[int, int]:
    """[lo, hi) of the utterances rank ``rank`` owns; blocks differ by at most one utterance."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def gather_utterance_values(local: torch.Tensor, total: int, group=None) -> torch.Tensor:
    """All-gather per-utterance rows ``[n_local, ...]`` into global utterance order ``[total, ...]`` on every rank.

    Blocks may be ragged (``total`` not divisible by the world size): rows are padded to the largest block for the
    collective and trimmed afterwards.  Works on any backend (NCCL on GPUs, gloo in the CPU tests).
    """
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = [shard_bounds(total, world, r) for r in range(world)]
    assert local.shape[0] == sizes[rank][1] - sizes[rank][0], "local block does not match this rank's shard"
    cap = max(hi - lo for lo, hi in sizes)
    pad = local.new_zeros((cap,) + tuple(local.shape[1:]))
    pad[: local.shape[0]] = local
    out: List[torch.Tensor] = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad, group=group)
    return torch.cat([out[r][: sizes[r][1] - sizes[r][0]] for r in range(world)], dim=0)
