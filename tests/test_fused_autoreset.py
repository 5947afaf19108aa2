"""
The FUSED auto-reset: ``step_random(auto_reset=True)`` = ``prepare()`` + ``step_random()`` in one launch (plus the scan of the
finished-game counts the previous step published).  It must leave EVERY array exactly as the two separate calls do -- both
ping-pong boards and masks (``prev_state`` is the post-reset board), ids (slot-ordered, game_numba.py:641-644), the ascending
index list, step/score/reward/flags, ``merged``, the one-hot rows -- for every kernel instance (64 / 256 / 768-thread blocks),
both spawn modes, eager and scheduled (CUDA graph), and however fused and plain calls are interleaved.  One case is held to the
CPU oracle directly.
"""

from __future__ import annotations

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

STATE = ("_board", "_valid", "_id", "_step_score", "_reward", "_terminated_padded", "_invalid", "_reset_count_dev", "_game_count_dev",
         "_stats_dev")


@pytest.fixture(scope="module")
def ml():
    import ml2048_b200

    assert torch.cuda.is_available()
    return ml2048_b200


def assert_same_state(a, b, what: str) -> None:
    assert a._cur == b._cur, what
    for name in STATE:
        assert torch.equal(getattr(a, name), getattr(b, name)), f"{name} {what}"
    if a._merged is not None:
        assert torch.equal(a._merged, b._merged), f"_merged {what}"
    if a._onehot is not None:
        assert torch.equal(a._onehot, b._onehot), f"_onehot {what}"
    n = int(a._reset_count_dev.item())
    assert torch.equal(a._reset_indices_dev[:n], b._reset_indices_dev[:n]), f"reset indices {what}"


@pytest.mark.parametrize("rng_mode", ["replay", "philox"])
@pytest.mark.parametrize("m,onehot,merged,steps", [
    (1, None, True, 400), (31, "u8", True, 300), (33, None, False, 300), (1000, "f32", True, 260),
    (8193, None, True, 200), (50021, "bf16", False, 180), ((1 << 19) + 5, None, False, 150), ((1 << 19) + 5, "f32", False, 150),
])
def test_fused_equals_prepare_then_step(ml, rng_mode, m, onehot, merged, steps):
    kw = dict(rng_mode=rng_mode, output="torch", onehot=onehot, track_merged=merged, sync_free=True)
    plain = ml.VecGame(m, "improved", **kw)
    fused = ml.VecGame(m, "improved", **kw)
    plain.reset(7)
    fused.reset(7)
    resets = 0
    for t in range(steps):
        plain.prepare()
        plain.step_random(return_actions=True)
        fused.step_random(return_actions=True, auto_reset=True)
        if t < 3 or t % 25 == 0 or t == steps - 1:
            assert_same_state(plain, fused, f"after step {t}")
            assert torch.equal(plain.sampled_actions, fused.sampled_actions), t
        resets += int(fused._reset_count_dev.item()) if t % 10 == 0 else 0
    assert plain._game_count == fused._game_count > m
    (i0,) = fused.last_reset()
    assert i0.dtype == torch.int64 and bool((i0[1:] > i0[:-1]).all())


def test_fused_steps_against_the_oracle(ml, oracle):
    """The fused path directly against the CPU oracle: the actions the kernel chose are replayed through
    prepare() + step(actions) of the oracle; every field, the ids and the reset index lists must agree."""
    m, n = 70001, 220
    env = ml.VecGame(m, "improved", output="torch", sync_free=True)
    env.reset(11)
    ref = oracle.OracleVecGame(m, "improved")
    ref.reset(11)
    for t in range(n):
        res = env.step_random(return_actions=True, auto_reset=True)
        (want_idx,) = ref.prepare()
        acts = env.sampled_actions.cpu().numpy().astype(np.int64)
        want = ref.step(acts)
        if t < 3 or t % 20 == 0 or t == n - 1:
            (got_idx,) = env.last_reset()
            np.testing.assert_array_equal(got_idx.cpu().numpy(), want_idx, err_msg=f"reset indices step {t}")
            for k in ("state", "valid_actions", "merged", "step", "terminated", "invalid", "prev_state", "prev_valid_actions"):
                np.testing.assert_array_equal(res[k].cpu().numpy(), want[k], err_msg=f"{k} step {t}")
            for k in ("reward", "score"):
                np.testing.assert_array_equal(res[k].cpu().numpy().view(np.uint32), want[k].view(np.uint32), err_msg=f"{k} step {t}")
            np.testing.assert_array_equal(env._id.cpu().numpy(), ref._data["id"], err_msg=f"id step {t}")
    assert env._game_count == ref._game_count > 2 * m


def test_interleaving_fused_and_plain_calls(ml):
    """Plain prepare()/step(), reset() and snapshots invalidate the published counts; the next fused step recounts."""
    m = 20000
    a = ml.VecGame(m, output="torch", sync_free=True)
    b = ml.VecGame(m, output="torch", sync_free=True)
    a.reset(3)
    b.reset(3)
    snap = None
    for t in range(240):
        a.prepare()
        a.step_random()
        if (t // 7) % 2 == 0:
            b.step_random(auto_reset=True)
        else:
            b.prepare()
            b.step_random()
        if t == 120:
            snap = (a.state_dict(), b.state_dict())
        if t % 30 == 0:
            for name in ("_board", "_valid", "_id", "_step_score", "_reward", "_terminated_padded", "_invalid", "_game_count_dev"):
                assert torch.equal(getattr(a, name), getattr(b, name)), (name, t)
    a.load_state_dict(snap[0])
    b.load_state_dict(snap[1])
    for t in range(40):
        a.prepare()
        a.step_random()
        b.step_random(auto_reset=True)
    assert_same_state(a, b, "after restore")
    a.reset(5)
    b.reset(5)
    for t in range(30):
        a.prepare()
        a.step_random()
        b.step_random(auto_reset=True)
    assert_same_state(a, b, "after reset()")


@pytest.mark.parametrize("m,onehot", [(2048, "f32"), (100000, None)])
def test_graphed_fused_rollout_equals_eager(ml, m, onehot):
    steps, replays = 16, 12
    eager = ml.VecGame(m, "improved", output="torch", onehot=onehot, sync_free=True)
    eager.reset(5)
    for _ in range(steps * replays):
        eager.prepare()
        eager.step_random()
    env = ml.VecGame(m, "improved", output="torch", onehot=onehot, sync_free=True)
    env.reset(5)
    roll = ml.GraphedRollout(env, steps, window=steps * 5, auto_reset=True)
    roll.replay(replays)
    torch.cuda.synchronize()
    assert_same_state(eager, env, "graph replay")


def test_fused_reset_argument_errors(ml):
    env = ml.VecGame(100, output="torch", sync_free=True)
    env.reset(1)
    env.enable_episode_log(10)
    with pytest.raises(RuntimeError):
        env.step_random(auto_reset=True)
