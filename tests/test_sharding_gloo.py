"""world_size-2 gloo tests (CPU) of the N>1 host logic: shard bounds, id offsets, statistics reduction."""

from __future__ import annotations

import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_everything():
    from ml2048_b200.sharding import shard_bounds

    for total in (1, 7, 8, 1000, 2**27):
        for world in (1, 2, 3, 8):
            if total < world:
                continue
            spans = [shard_bounds(total, world, r) for r in range(world)]
            assert spans[0][0] == 0
            for (s0, n0), (s1, _) in zip(spans, spans[1:]):
                assert s0 + n0 == s1
            assert spans[-1][0] + spans[-1][1] == total
            assert max(n for _, n in spans) - min(n for _, n in spans) <= 1
    with pytest.raises(ValueError):
        shard_bounds(0, 1, 0)


def _worker(rank: int, world: int, port: int, out_dir: str) -> None:
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from ml2048_b200.sharding import exclusive_id_offset, reduce_episode_stats, reduce_live_histogram, shard_bounds

    # each rank "plays" its shard of a deterministic synthetic episode log
    total = 1001
    base, size = shard_bounds(total, world, rank)
    rng = np.random.default_rng(0)
    max_tile = rng.integers(3, 10, size=total)
    score = rng.integers(100, 5000, size=total)
    steps = rng.integers(50, 300, size=total)
    sl = slice(base, base + size)
    stats = torch.zeros(24, dtype=torch.int64)
    stats[:20] = torch.from_numpy(np.bincount(max_tile[sl], minlength=20))
    stats[20] = size
    stats[21] = int(score[sl].sum())
    stats[22] = int(steps[sl].sum())
    stats[23] = int(score[sl].max())
    red = reduce_episode_stats(stats, None)
    want = torch.zeros(24, dtype=torch.int64)
    want[:20] = torch.from_numpy(np.bincount(max_tile, minlength=20))
    want[20], want[21], want[22], want[23] = total, int(score.sum()), int(steps.sum()), int(score.max())
    assert torch.equal(red, want), (rank, red, want)
    assert torch.equal(reduce_live_histogram(stats[:20].clone(), None), want[:20])

    # id offsets: all_gather of per-rank reset counts -> exclusive prefix
    mine = torch.tensor([3 + 5 * rank], dtype=torch.int64)
    gathered = [torch.zeros(1, dtype=torch.int64) for _ in range(world)]
    dist.all_gather(gathered, mine)
    counts_all = torch.cat(gathered)
    off = int(exclusive_id_offset(counts_all, rank))
    assert off == sum(3 + 5 * r for r in range(rank))
    with open(os.path.join(out_dir, f"ok{rank}"), "w") as fh:
        fh.write("ok")
    dist.destroy_process_group()


def test_stats_reduction_world2_gloo(tmp_path):
    world = 2
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, str(tmp_path)), nprocs=world, join=True)
    assert all((tmp_path / f"ok{r}").exists() for r in range(world))


def test_reduce_is_identity_without_process_group():
    from ml2048_b200.sharding import reduce_episode_stats

    s = torch.arange(24, dtype=torch.int64)
    assert torch.equal(reduce_episode_stats(s), s)
    with pytest.raises(ValueError):
        reduce_episode_stats(torch.zeros(5, dtype=torch.int64))
