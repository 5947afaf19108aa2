"""Multi-GPU sharding of the environment: one process per GPU, games split into contiguous slot ranges.

Games are independent (each iteration of the reference's ``prange`` touches only its own record,
reference: src/ml2048/game_numba.py:715-738), so the data path needs NO collective.  What is exchanged:

  * finished-episode statistics -- the max-tile histogram of ``RunnerStats`` (runner.py:150-189, whose
    ``combine`` sums histograms) plus episode/score/step sums (all-reduce SUM) and the max score
    (all-reduce MAX): 24 integers, every K steps;
  * optionally the per-rank reset counts of one ``prepare()`` (all-gather of one integer per rank) so
    that game ids stay globally slot-ordered like in the single-process reference (game_numba.py:641-644).

The helpers below work on CPU tensors with the gloo backend (tests) and on CUDA tensors with NCCL.
"""

from __future__ import annotations

from typing import Any, Optional

import torch
import torch.distributed as dist

STATS_SUM_WORDS = 23  # hist[20], episodes, score_sum, step_sum  -> SUM;  word 23 = score_max -> MAX


def shard_bounds(total_games: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous split of ``total_games`` global slots: returns (slot_base, size) of ``rank``.
    The first ``total % world`` ranks hold one extra game."""
    if total_games <= 0 or world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError((total_games, world_size, rank))
    base, extra = divmod(total_games, world_size)
    size = base + (1 if rank < extra else 0)
    start = rank * base + min(rank, extra)
    return start, size


def exclusive_id_offset(counts_all: torch.Tensor, rank: int) -> torch.Tensor:
    """Number of games reset on lower ranks in this prepare(): the id offset of this shard."""
    return counts_all[:rank].sum()


def reduce_episode_stats(stats: torch.Tensor, group: Optional[Any] = None) -> torch.Tensor:
    """All-reduce the int64[24] statistics vector of ``VecGame.episode_stats_tensor()`` over ``group``.
    SUM for the histogram and the sums, MAX for the max score.  Returns a new tensor (same device)."""
    if stats.shape != (24,) or stats.dtype != torch.int64:
        raise ValueError(f"expected int64[24], got {stats.dtype}{tuple(stats.shape)}")
    sums = stats[:STATS_SUM_WORDS].clone()
    mx = stats[STATS_SUM_WORDS:].clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(sums, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(mx, op=dist.ReduceOp.MAX, group=group)
    return torch.cat([sums, mx])


def reduce_live_histogram(hist20: torch.Tensor, group: Optional[Any] = None) -> torch.Tensor:
    """All-reduce (SUM) of the live-board max-tile histogram behind ``VecGame.summary()``
    (game_numba.py:593-604)."""
    out = hist20.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out
