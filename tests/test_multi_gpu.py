"""N>1 on real GPUs (skipped with fewer than 2): a 2-rank NCCL run, sharded by global slot with
globally slot-ordered ids, must equal the single-process run slot for slot -- boards, scores, ids -- and
the all-reduced episode statistics must equal the single-process statistics."""

from __future__ import annotations

import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M_TOTAL, STEPS, SEED = 6000, 260, 17


def _worker(rank: int, world: int, port: int, out_dir: str) -> None:
    sys.path.insert(0, ROOT)
    import torch.distributed as dist

    import ml2048_b200
    from ml2048_b200.sharding import reduce_episode_stats, shard_bounds

    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    base, size = shard_bounds(M_TOTAL, world, rank)
    env = ml2048_b200.VecGame(size, output="torch", slot_base=base, sync_free=True)
    env.shard()
    env.reset(SEED)
    for _ in range(STEPS):
        env.prepare()
        env.step_random()
    stats = reduce_episode_stats(env.episode_stats_tensor())
    torch.save({"board": env.observations()[0].cpu(), "id": env._id.cpu(), "score": env._score.cpu(),
                "stats": stats.cpu(), "game_count": env._game_count}, os.path.join(out_dir, f"rank{rank}.pt"))
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_run_equals_single_process(tmp_path):
    import torch.multiprocessing as mp

    import ml2048_b200

    world = 2
    mp.spawn(_worker, args=(world, 29650 + os.getpid() % 300, str(tmp_path)), nprocs=world, join=True)
    parts = [torch.load(tmp_path / f"rank{r}.pt") for r in range(world)]
    whole = ml2048_b200.VecGame(M_TOTAL, output="torch", sync_free=True)
    whole.reset(SEED)
    for _ in range(STEPS):
        whole.prepare()
        whole.step_random()
    assert torch.equal(torch.cat([p["board"] for p in parts]), whole.observations()[0].cpu())
    assert torch.equal(torch.cat([p["score"] for p in parts]), whole._score.cpu())
    assert torch.equal(torch.cat([p["id"] for p in parts]), whole._id.cpu())
    assert parts[0]["game_count"] == parts[1]["game_count"] == whole._game_count
    assert torch.equal(parts[0]["stats"], whole.episode_stats_tensor().cpu())
    assert torch.equal(parts[0]["stats"], parts[1]["stats"])
    assert int(parts[0]["stats"][20]) > 5000
