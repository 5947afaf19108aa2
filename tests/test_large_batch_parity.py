"""
GPU parity of the LARGE-BATCH kernel instances -- the ones the benchmark is quoted on:

    step_kernel<rng, log, onehot, full, 768>   M >= 2^19 with a fused one-hot: 768-thread blocks, each writing its
                                               one-hot tile through the TMA bulk-copy engine in 48 KiB chunks
                                               (write_onehot_tile_tma: double buffer, wait_group.read, tail chunk)
    prepare_fused_kernel + write_onehot_listed the cooperative single-launch auto-reset and its one-hot rows

against the CPU oracle (oracle/, pinned to the live reference by tests/golden/) in lock step: all ten VecStepResult
fields (game_numba.py:507-519) + ids + reset indices, and the one-hot compared element for element with the
reference encoding ``F.one_hot(x, 16).float().permute(0, 2, 1)`` (policy/_network.py:86-95) AFTER prepare() (reset
rows) AND AFTER step().  Sizes: exact multiples of the 768-game block, sizes whose last block ends in a partial
(tail) chunk, a last block of ONE game, and the benchmark size M = 2^24 (fp32, compared in slices).

The environments start from a STEADY STATE transplanted from a fast in-kernel burn-in (about 0.9 % of the games end
per step there), so every compared step exercises auto-resets, not only the all-games reset of the first prepare().
"""

from __future__ import annotations

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

BLOCK = 768  # games per block of the large-batch instance (ML2048_ONEHOT_STEP_THREADS)
TORCH_DTYPE = {"f32": torch.float32, "bf16": torch.bfloat16, "u8": torch.uint8}


@pytest.fixture(scope="module")
def ml():
    import ml2048_b200

    assert torch.cuda.is_available()
    return ml2048_b200


def reference_onehot(board: torch.Tensor) -> torch.Tensor:
    """policy/_network.py:86-95 on a (n,16) uint8 CUDA tensor."""
    return torch.nn.functional.one_hot(board.long(), 16).float().permute(0, 2, 1)


def assert_onehot_equal(env, what: str, slice_games: int = 1 << 18) -> None:
    board = env.observations()[0]
    oh = env.observations_onehot()
    assert oh.shape == (env._size, 16, 16)
    for lo in range(0, env._size, slice_games):
        hi = min(env._size, lo + slice_games)
        want = reference_onehot(board[lo:hi])
        got = oh[lo:hi].float()
        if not torch.equal(got, want):
            bad = (got != want).flatten(1).any(dim=1).nonzero().flatten()
            raise AssertionError(f"one-hot {what}: {bad.numel()} games differ in [{lo},{hi}), first game {lo + int(bad[0])} "
                                 f"(block {(lo + int(bad[0])) // BLOCK}, game-in-block {(lo + int(bad[0])) % BLOCK})")


def steady_state(ml, m: int, seed: int, steps: int = 160, **kw) -> dict[str, np.ndarray]:
    """Boards in steady state (finished and running games mixed) from a fast device-side rollout."""
    env = ml.VecGame(m, "improved", output="torch", sync_free=True, **kw)
    env.reset(seed)
    for _ in range(steps):
        env.prepare()
        env.step_random()
    torch.cuda.synchronize()
    cur = env._cur
    st = {"board": env._board[cur], "valid_actions": env._valid[cur], "id": env._id, "step": env._step, "score": env._score,
          "reward": env._reward, "terminated": env._terminated, "invalid": env._invalid, "merged": env._merged}
    out = {k: v.cpu().numpy().copy() for k, v in st.items()}
    out["game_count"] = env._game_count
    del env
    torch.cuda.empty_cache()
    return out


def transplant(env, ref, st: dict[str, np.ndarray]) -> None:
    """Load the same mid-rollout state into the CUDA environment and the oracle (both freshly reset with the same seed, so
    their host random schedules are in step)."""
    cur = env._cur
    dev = env.device
    env._board[cur].copy_(torch.from_numpy(st["board"]).to(dev))
    env._valid[cur].copy_(torch.from_numpy(st["valid_actions"]).to(dev))
    for name in ("id", "step", "score", "reward", "terminated", "invalid", "merged"):
        getattr(env, "_" + name).copy_(torch.from_numpy(st[name]).to(dev))
    env._game_count = int(st["game_count"])
    if env._onehot is not None:
        from ml2048_b200 import _lib

        _lib.check(env._lib.ml2048_encode_onehot(env._board[cur].data_ptr(), env._onehot.data_ptr(), env._onehot_kind, env._size,
                                                 torch.cuda.current_stream(dev).cuda_stream), "ml2048_encode_onehot")
    d = ref._data
    for name in ("board", "valid_actions", "id", "step", "score", "reward", "terminated", "invalid", "merged"):
        d[name] = st[name]
    ref._game_count = int(st["game_count"])


def assert_fields_equal(env, ref, res, want, t: int) -> None:
    dev = env.device
    for k in ("state", "valid_actions", "merged", "step", "terminated", "invalid", "prev_state", "prev_valid_actions"):
        assert torch.equal(res[k], torch.from_numpy(np.ascontiguousarray(want[k])).to(dev)), f"{k} differs at step {t}"
    for k in ("reward", "score"):
        w = torch.from_numpy(np.ascontiguousarray(want[k]).view(np.int32)).to(dev)
        assert torch.equal(res[k].contiguous().view(torch.int32), w), f"{k} differs (bitwise) at step {t}"
    assert torch.equal(env._id, torch.from_numpy(np.ascontiguousarray(ref._data["id"])).to(dev)), f"id differs at step {t}"


def lockstep(ml, oracle, m: int, onehot: str, steps: int, seed: int, reward: str = "improved", two_prob: float = 0.8,
             onehot_slice: int = 1 << 18) -> int:
    st = steady_state(ml, m, seed)
    ref = oracle.OracleVecGame(m, reward, two_prob=two_prob)
    ref.reset(seed + 1)
    env = ml.VecGame(m, reward, two_prob=two_prob, output="torch", onehot=onehot)
    env.reset(seed + 1)
    transplant(env, ref, st)
    assert env.observations_onehot().dtype == TORCH_DTYPE[onehot]
    acts = np.empty((m,), np.int64)
    resets = 0
    for t in range(steps):
        (i0,) = ref.prepare()
        (i1,) = env.prepare()
        assert torch.equal(i1, torch.from_numpy(i0).to(env.device)), f"reset indices differ at step {t}"
        resets += i0.size
        assert_onehot_equal(env, f"after prepare() {t}", onehot_slice)  # write_onehot_listed rows + untouched rows
        ref.random_valid_actions(1000 * seed + t, acts)
        if t % 3 == 1:
            acts[::11] = (acts[::11] + 1) % 4  # some wrong directions: the invalid-move path
        want = ref.step(acts)
        res = env.step(torch.from_numpy(acts).to(env.device))
        assert_fields_equal(env, ref, res, want, t)
        assert_onehot_equal(env, f"after step() {t}", onehot_slice)  # the TMA tile writer
    assert env._game_count == ref._game_count
    return resets


SIZES = [
    1 << 19,             # smallest batch that selects the 768-thread instance; last block 512 games
    (1 << 19) + 5,       # last block 517 games: f32 tail chunk of 37 games (48 per chunk), bf16 37 of 96, u8 133 of 192
    BLOCK * 700,         # exact multiple of the block: no partial block
    BLOCK * 683 + 1,     # last block holds ONE game
    (1 << 20) + 3,       # last block 259 games: tail chunks of 19 (f32), 67 (bf16, u8)
]


@pytest.mark.parametrize("onehot", ["f32", "bf16", "u8"])
@pytest.mark.parametrize("m", SIZES)
def test_large_batch_replay_lockstep_with_onehot(ml, oracle, m, onehot):
    resets = lockstep(ml, oracle, m, onehot, steps=6, seed=m % 97 + 3)
    assert resets > 6 * m // 400, "the steady state must produce auto-resets in every compared step"


@pytest.mark.parametrize("reward,two_prob", [("normal", 0.8), ("rank", 0.3), ("maxcell", 1.0)])
def test_large_batch_replay_other_rewards(ml, oracle, reward, two_prob):
    lockstep(ml, oracle, (1 << 19) + 777, "f32", steps=4, seed=17, reward=reward, two_prob=two_prob)


def test_bench_size_fp32_lockstep(ml, oracle):
    """M = 2^24, fp32 one-hot, replay tables: the configuration the headline number is quoted on (17 GB of one-hot rows,
    compared in slices of 2^20 games), every field against the oracle."""
    lockstep(ml, oracle, 1 << 24, "f32", steps=3, seed=5, onehot_slice=1 << 20)


def test_extended_validation_config(ml, oracle):
    """One configuration of tools/extended_validation.py inside the suite: M = 2^20, 150 runner steps from reset(), every state
    field compared every 25 steps, the fused u8 one-hot compared with the reference encoding (not only sum == 1)."""
    m, n = 1 << 20, 150
    ref = oracle.OracleVecGame(m, "rank", two_prob=0.5)
    ref.reset(3)
    env = ml.VecGame(m, "rank", two_prob=0.5, output="torch", onehot="u8")
    env.reset(3)
    acts = np.empty(m, np.int64)
    for t in range(n):
        (i0,) = ref.prepare()
        (i1,) = env.prepare()
        assert i1.numel() == i0.size, t
        ref.random_valid_actions(300000 + t, acts)
        if t % 7 == 3:
            acts[::13] = (acts[::13] + 1) % 4
        want = ref.step(acts)
        res = env.step(torch.from_numpy(acts).cuda())
        if t % 25 == 24 or t == n - 1:
            assert torch.equal(i1, torch.from_numpy(i0).cuda())
            assert_fields_equal(env, ref, res, want, t)
            assert_onehot_equal(env, f"after step() {t}")
    assert env._game_count == ref._game_count > m


# ---------------------------------------------------------------------------------------------
# Philox mode: no CPU counterpart of the spawn stream, so the large-batch instance is held to the small ones
# ---------------------------------------------------------------------------------------------


@pytest.mark.parametrize("onehot", ["f32", "bf16", "u8"])
def test_large_batch_philox_equals_small_batch_instances(ml, onehot):
    """Philox draws are keyed by (seed, global slot, step counter), so the SAME games played as one large batch
    (768-thread TMA instance, cooperative auto-reset) and as shards small enough to select the 256-thread and the
    64-thread plain-store instances (and the single-block / three-launch auto-resets) must agree bit for bit: boards,
    masks, scores, rewards, flags and the one-hot rows, which are also held to the reference encoding."""
    m = (1 << 19) + 5
    cuts = [0, 30000, 30000 + (1 << 18), m]  # 64-thread instance, 256-thread instance, 256-thread instance
    whole = ml.VecGame(m, "improved", rng_mode="philox", output="torch", onehot=onehot, sync_free=True)
    parts = [ml.VecGame(hi - lo, "improved", rng_mode="philox", output="torch", onehot=onehot, sync_free=True, slot_base=lo)
             for lo, hi in zip(cuts[:-1], cuts[1:])]
    for e in [whole] + parts:
        e.reset(123)
    for t in range(140):
        for e in [whole] + parts:
            e.prepare()
        if t in (0, 1, 100, 139):
            assert_onehot_equal(whole, f"after prepare() {t}")
            assert torch.equal(torch.cat([p.observations_onehot() for p in parts]), whole.observations_onehot()), f"prepare {t}"
        for e in [whole] + parts:
            e.step_random()
        if t in (0, 1, 100, 139):
            assert_onehot_equal(whole, f"after step() {t}")
            assert torch.equal(torch.cat([p.observations_onehot() for p in parts]), whole.observations_onehot()), f"step {t}"
    for name in ("_score", "_step", "_reward", "_terminated", "_invalid", "_merged"):
        assert torch.equal(torch.cat([getattr(p, name) for p in parts]), getattr(whole, name)), name
    for k in (0, 1):
        assert torch.equal(torch.cat([p.observations()[k] for p in parts]), whole.observations()[k])
    st = whole.episode_stats()
    assert st["episodes"] > m // 4
    assert st["episodes"] == sum(p.episode_stats()["episodes"] for p in parts)
