"""
Pins the CPU oracle (oracle/vecgame_oracle.c) against fixtures generated from the live reference
(oracle/gen_golden.py).  CPU only.  Reference: src/ml2048/game_numba.py.
"""

from __future__ import annotations

import numpy as np
import pytest

from conftest import golden, golden_rollouts
from oracle.rollout import compare_rollouts, record_rollout


def test_known_answer_playground(oracle):
    # the only known-answer vector the reference holds: playground.ipynb:3915-3926, output :3901-3906
    prev = np.array([10, 10, 8, 10, 10, 9, 8, 10, 9, 10, 9, 9, 10, 10, 9, 9], np.uint8)
    want = np.array([11, 8, 10, 0, 10, 9, 8, 10, 9, 10, 10, 0, 11, 10, 0, 0], np.uint8)
    state, merged = oracle.board_move(prev, 0)
    np.testing.assert_array_equal(state, want)
    want_merged = np.zeros(16, np.uint8)
    want_merged[9] = 2
    want_merged[10] = 2
    np.testing.assert_array_equal(merged, want_merged)
    assert oracle.board_reward("normal", state, prev, merged) == 6144.0
    assert oracle.board_reward("rank", state, prev, merged) == 42.0
    assert oracle.board_reward("maxcell", state, prev, merged) == 2052.0
    kat = golden("kat_playground.npz")
    np.testing.assert_array_equal(kat["state"].astype(np.uint8), state)
    np.testing.assert_array_equal(kat["merged"].astype(np.uint8), merged)


def test_push_examples(oracle):
    # SURVEY.md section 4 examples probed on the reference's _push_row (game_numba.py:48-90)
    for src, dst in (([1, 1, 1, 1], [2, 2, 0, 0]), ([1, 1, 1, 0], [2, 1, 0, 0]), ([2, 1, 1, 2], [2, 2, 2, 0]),
                     ([0, 1, 0, 1], [2, 0, 0, 0])):
        got, _ = oracle.line_push(np.array(src, np.uint8), False)
        assert got.tolist() == dst


def test_line_table_exhaustive(oracle):
    tab = golden("line_table.npz")
    v = np.arange(18, dtype=np.uint8)
    lines = np.stack(np.meshgrid(v, v, v, v, indexing="ij"), axis=-1).reshape(-1, 4)
    for name, toward_last in (("first", False), ("last", True)):
        pushed = tab[f"pushed_{name}"]
        fused = tab[f"fused_{name}"]
        for i in range(0, lines.shape[0], 1):
            got, buckets = oracle.line_push(lines[i], toward_last)
            if not np.array_equal(got, pushed[i]):
                raise AssertionError((lines[i], got, pushed[i]))
            ks = np.repeat(np.arange(18), buckets)
            want = fused[i][fused[i] > 0]
            if not np.array_equal(ks, want):
                raise AssertionError((lines[i], ks, want))
    mf, ml = tab["movable_first"], tab["movable_last"]
    for i in range(lines.shape[0]):
        f, b = oracle.line_flags(*[int(x) for x in lines[i]])
        assert f == mf[i] and b == ml[i], lines[i]
        # property probed in the survey: movable == "pushing the line changes it"
        assert f == (not np.array_equal(tab["pushed_first"][i], lines[i]))
        assert b == (not np.array_equal(tab["pushed_last"][i], lines[i]))


def test_boards_moves_masks_rewards(oracle):
    g = golden("boards.npz")
    names = [str(x) for x in g["reward_names"]]
    boards = g["boards"]
    for i in range(boards.shape[0]):
        np.testing.assert_array_equal(oracle.board_valid(boards[i]), g["mask"][i])
        for a in range(4):
            moved, merged = oracle.board_move(boards[i], a)
            np.testing.assert_array_equal(moved, g["moved"][i, a])
            np.testing.assert_array_equal(merged, g["merged"][i, a])
            for j, nm in enumerate(names):
                assert oracle.board_reward(nm, moved, boards[i], merged) == g["rewards"][i, a, j]


def test_onehot_layout(oracle):
    # policy/_network.py:86-95: one_hot(x,16).float().permute(0,2,1)
    import torch
    import torch.nn.functional as F

    boards = golden("boards.npz")["boards"]
    want = F.one_hot(torch.from_numpy(boards).long(), 16).float().permute(0, 2, 1).contiguous().numpy()
    np.testing.assert_array_equal(oracle.onehot(boards), want)


@pytest.mark.parametrize("fname", golden_rollouts())
def test_rollout_matches_reference(oracle, fname):
    g = golden(fname)
    m, n, seed, aseed = [int(x) for x in g["meta"]]
    two_prob, wild = [float(x) for x in g["meta_f"]]
    env = oracle.OracleVecGame(m, str(g["reward_kind"]), two_prob=two_prob)
    env.reset(seed)
    got = record_rollout(env, n, action_seed=aseed, wild=wild, full=True)
    compare_rollouts(got, g)


def test_recorded_schedule_replay(oracle):
    g = golden("schedule_sched_small.npz")
    m, n, seed, aseed = [int(x) for x in g["meta"]]
    two_prob, wild = [float(x) for x in g["meta_f"]]
    sched = oracle.RecordedSchedule(g["sched_coins"], g["sched_offsets"], g["sched_perms"], g["sched_floats"])
    env = oracle.OracleVecGame(m, str(g["reward_kind"]), two_prob=two_prob)
    env.reset(schedule=sched)
    got = record_rollout(env, n, actions=g["actions"], full=True)
    compare_rollouts(got, g)


def test_invalid_move_keeps_stale_fields(oracle):
    # game_numba.py:737-738: an invalid action only sets invalid = 1
    env = oracle.OracleVecGame(64)
    env.reset(3)
    rng = np.random.default_rng(0)
    for _ in range(30):
        env.prepare()
        _, valid = env.observations()
        acts = oracle.random_valid_actions(valid, rng.random(64))
        env.step(acts)
    env.prepare()
    valid = env.observations()[1].copy()  # observations() are views into the live records
    before = env._data.copy()
    bad = np.array([int(np.argmin(v)) for v in valid], dtype=np.int64)  # first invalid direction if any
    res = env.step(bad)
    is_invalid = valid[np.arange(64), bad] == 0
    assert is_invalid.any()
    np.testing.assert_array_equal(res["invalid"].astype(bool), is_invalid)
    for f in ("board", "merged", "step", "score", "reward", "terminated", "valid_actions"):
        np.testing.assert_array_equal(env._data[f][is_invalid], before[f][is_invalid])


def test_ctor_errors(oracle):
    with pytest.raises(ValueError):
        oracle.OracleVecGame(0)
    env = oracle.OracleVecGame(4)
    with pytest.raises(AssertionError):
        env.step(np.zeros(5, np.int64))


def test_runner_stats_fixture_from_the_reference_runner(oracle):
    """RunnerStats of the reference's own VecRunner (tests/golden/runner_stack.npz) == the oracle's _update_count
    restatement (orc_terminated_hist, runner.py:120-136) under the same deterministic policy; also the per-id episode
    records the reference's ReplayRecorder produced (steps, max tile, score)."""
    g = golden("runner_stack.npz")
    m, steps, seed = [int(x) for x in g["meta"]]
    env = oracle.OracleVecGame(m, "improved")
    env.reset(seed)
    counts = np.zeros(20, np.int64)
    total = 0
    by_id = {}
    for t in range(steps):
        env.prepare()
        valid = env.observations()[1].astype(bool)
        nvalid = valid.sum(axis=1)
        k = (7 * t + 13 * np.arange(m)) % np.maximum(nvalid, 1)
        rank = np.cumsum(valid, axis=1) - 1
        acts = np.where(nvalid > 0, (valid & (rank == k[:, None])).argmax(axis=1), 0).astype(np.int64)
        ids = env._data["id"].copy()
        res = env.step(acts)
        c, n = env.terminated_hist()
        counts += c
        total += n
        for slot in np.flatnonzero(res["terminated"]):
            by_id[int(ids[slot])] = (int(res["step"][slot]), int(res["state"][slot].max()), float(res["score"][slot]))
    np.testing.assert_array_equal(counts, g["stats_counts"])
    assert total == int(g["stats_terminated"]) and env._game_count == int(g["game_count"])
    for i, gid in enumerate(g["buf_id"].tolist()):
        assert by_id[gid] == (int(g["buf_steps"][i]), int(g["buf_maxcell"][i]), float(g["buf_score"][i])), gid
