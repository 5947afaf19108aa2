"""
oracle/gen_golden.py -- generate tests/golden/*.npz from the LIVE reference.  TEST INFRASTRUCTURE.

Runs only in the authoring container, where the unmodified reference is mounted read-only at
/root/reference (Python + Numba; it cannot travel to the GPU box, hence committed fixtures):

    PYTHONDONTWRITEBYTECODE=1 python -m oracle.gen_golden

What is produced (all from the reference's own functions, reference: src/ml2048/game_numba.py):

  kat_playground.npz   the one known-answer vector the reference holds (playground.ipynb:3915-3926,
                       printed output :3901-3906), re-run through the live reference here
  line_table.npz       exhaustive 18^4 lines x {toward cell 0, toward cell 3}: pushed line, fused
                       exponents (_push_row :48-90) and the movable flags (_line_movable :215-256)
  boards.npz           random boards x 4 actions: moved board, merged, mask, the four rewards
  rollout_*.npz        lock-step rollouts through VecGame (prepare/observations/step), every
                       VecStepResult field + ids + reset indices (full or CRC32 digests)
  runner_stack.npz     the reference's VecRunner + RunnerStats + ReplayRecorder driven by a deterministic policy
  gae.npz              compute_gae (gae.py:7-68) on random fp32 inputs with a stand-in critic
  random_policy_stats.npz  per-seed episode statistics of the reference under the uniform-over-valid policy (16 seeds)
  schedule_*.npz       the host random draws of one rollout recorded from the reference's generator
                       (tables at every refresh, coins, offsets): input of the "replay" mode
"""

from __future__ import annotations

import os
import sys

import numpy as np

os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
sys.dont_write_bytecode = True
REF_SRC = "/root/reference/src"
sys.path.insert(0, REF_SRC)

import numba  # noqa: E402
from ml2048 import game_numba as ref  # noqa: E402

from .rollout import record_rollout  # noqa: E402

OUT_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

REWARDS = {
    "normal": ref.reward_fn_normal,
    "improved": ref.reward_fn_improved,
    "rank": ref.reward_fn_rank,
    "maxcell": ref.reward_fn_maxcell,
}

# (name, M, N, env seed, reward, two_prob, action seed, wild fraction, keep full arrays)
ROLLOUTS = [
    ("c1_seed0_normal", 1024, 64, 0, "normal", 0.8, 7, 0.05, True),     # BASELINE config 1
    ("c1_seed1_improved", 1024, 64, 1, "improved", 0.8, 7, 0.05, False),
    ("c1_seed123_normal", 1024, 64, 123, "normal", 0.8, 7, 0.0, False),
    ("ragged_rank", 1500, 300, 5, "rank", 0.8, 11, 0.05, False),          # M not a multiple of 1024, many resets
    ("twoprob_maxcell", 2304, 200, 9, "maxcell", 0.5, 13, 0.10, False),   # M > 1024: table-row aliasing
    ("tiny_long", 7, 1500, 3, "improved", 0.8, 17, 0.02, False),          # many refreshes, tiny M
    ("train_shape", 2048, 64, 2024, "improved", 0.8, 19, 0.0, False),     # BASELINE config 2
    ("one_game", 1, 400, 4, "normal", 0.8, 23, 0.05, False),
]

SCHEDULES = [("sched_small", 96, 80, 42, "improved", 0.8, 29, 0.05)]


class _RecordingGenerator:
    """Wraps the reference instance's numpy Generator and logs what VecGame draws from it."""

    def __init__(self, inner: np.random.Generator):
        self._inner = inner
        self.coins: list[float] = []
        self.offsets: list[int] = []
        self.perms: list[np.ndarray] = []
        self.floats: list[np.ndarray] = []

    def random(self, *args, **kwargs):
        r = self._inner.random(*args, **kwargs)
        if "out" in kwargs:
            self.floats.append(np.array(kwargs["out"], copy=True))
        else:
            self.coins.append(float(r))
        return r

    def integers(self, *args, **kwargs):
        r = self._inner.integers(*args, **kwargs)
        self.offsets.append(int(r))
        return r

    def permuted(self, *args, **kwargs):
        r = self._inner.permuted(*args, **kwargs)
        self.perms.append(np.array(kwargs["out"], copy=True))
        return r


@numba.njit
def _lines_via_reference(lines, toward_last, out_lines, out_buckets, out_f, out_b):
    n = lines.shape[0]
    for i in range(n):
        row = lines[i].copy()
        buckets = np.zeros(18, dtype=np.uint8)
        if toward_last:
            ref._push_row(row, 3, -1, buckets)
        else:
            ref._push_row(row, 0, 1, buckets)
        out_lines[i, :] = row
        out_buckets[i, :] = buckets
        f, b = ref._line_movable(lines[i, 0], lines[i, 1], lines[i, 2], lines[i, 3])
        out_f[i] = f
        out_b[i] = b


def gen_kat() -> None:
    # playground.ipynb:3915-3926
    np.random.seed(12322)
    state = np.random.randint(8, 11, (16,), dtype=np.int8)
    prev = state.copy()
    merged = np.zeros_like(state)
    ref._step_left(state, merged)
    rn = ref.reward_fn_normal(state, prev, merged)
    rr = ref.reward_fn_rank(state, prev, merged)
    rm = ref.reward_fn_maxcell(state, prev, merged)
    # the printed output of the notebook, :3901-3906
    assert prev.tolist() == [10, 10, 8, 10, 10, 9, 8, 10, 9, 10, 9, 9, 10, 10, 9, 9]
    assert state.tolist() == [11, 8, 10, 0, 10, 9, 8, 10, 9, 10, 10, 0, 11, 10, 0, 0]
    assert (rn, rr, rm) == (6144.0, 42.0, 2052.0)
    np.savez_compressed(os.path.join(OUT_DIR, "kat_playground.npz"), prev=prev, state=state, merged=merged,
                        reward_normal=rn, reward_rank=rr, reward_maxcell=rm)


def gen_line_table() -> None:
    v = np.arange(18, dtype=np.uint8)
    lines = np.stack(np.meshgrid(v, v, v, v, indexing="ij"), axis=-1).reshape(-1, 4)
    n = lines.shape[0]
    res = {}
    for name, toward_last in (("first", False), ("last", True)):
        out_lines = np.zeros((n, 4), np.uint8)
        out_buckets = np.zeros((n, 18), np.uint8)
        f = np.zeros(n, np.bool_)
        b = np.zeros(n, np.bool_)
        _lines_via_reference(lines, toward_last, out_lines, out_buckets, f, b)
        res[f"pushed_{name}"] = out_lines
        # at most two fusions per line: store the fused exponents (0 = none), ascending
        fused = np.zeros((n, 2), np.uint8)
        for i in np.flatnonzero(out_buckets.sum(axis=1)):
            ks = np.repeat(np.arange(18), out_buckets[i])
            fused[i, : ks.size] = ks
        res[f"fused_{name}"] = fused
        res["movable_first"] = f
        res["movable_last"] = b
    np.savez_compressed(os.path.join(OUT_DIR, "line_table.npz"), **res)


def gen_boards() -> None:
    rng = np.random.default_rng(20481)
    n = 4096
    # mixture: sparse early boards, dense late boards, boards with large exponents, dead boards
    dens = rng.choice([0.2, 0.5, 0.8, 1.0], size=n)
    hi = rng.choice([3, 6, 11, 15], size=n)  # fusing exponent >= 16 would index merged[16+] (out of bounds in the reference)
    boards = (rng.integers(1, 18, size=(n, 16)) % hi[:, None] + 1).astype(np.uint8)
    boards[rng.random((n, 16)) >= dens[:, None]] = 0
    boards[0] = 0
    boards[1] = np.array([1, 2, 1, 2, 2, 1, 2, 1, 1, 2, 1, 2, 2, 1, 2, 1], np.uint8)  # dead board
    boards[2] = 15
    moved = np.zeros((n, 4, 16), np.uint8)
    merged = np.zeros((n, 4, 16), np.uint8)
    mask = np.zeros((n, 4), np.uint8)
    rewards = np.zeros((n, 4, 4), np.float64)
    for i in range(n):
        ref._compute_valid_actions(boards[i], mask[i])
        for a in range(4):
            b = boards[i].copy()
            m = np.zeros(16, np.uint8)
            ref._step_kernel(b, m, a)
            moved[i, a] = b
            merged[i, a] = m
            for j, fn in enumerate(REWARDS.values()):
                rewards[i, a, j] = fn(b, boards[i], m)
    np.savez_compressed(os.path.join(OUT_DIR, "boards.npz"), boards=boards, moved=moved, merged=merged, mask=mask,
                        rewards=rewards, reward_names=np.array(list(REWARDS)))


def gen_rollouts() -> None:
    for name, m, n, seed, reward, two_prob, aseed, wild, full in ROLLOUTS:
        vg = ref.VecGame(m, REWARDS[reward], two_prob=two_prob)
        vg.reset(seed)
        rec = record_rollout(vg, n, action_seed=aseed, wild=wild, full=full)
        rec["meta"] = np.array([m, n, seed, aseed], dtype=np.int64)
        rec["meta_f"] = np.array([two_prob, wild], dtype=np.float64)
        rec["reward_kind"] = np.array(reward)
        np.savez_compressed(os.path.join(OUT_DIR, f"rollout_{name}.npz"), **rec)
        print(f"rollout {name}: resets={int(rec['n_reset'].sum())} games={int(rec['final_game_count'])}")


def gen_schedules() -> None:
    for name, m, n, seed, reward, two_prob, aseed, wild in SCHEDULES:
        vg = ref.VecGame(m, REWARDS[reward], two_prob=two_prob)
        vg.reset(seed)
        first_perm, first_float = vg._randperm.copy(), vg._randfloat.copy()
        proxy = _RecordingGenerator(vg._rand)
        vg._rand = proxy
        rec = record_rollout(vg, n, action_seed=aseed, wild=wild, full=True)
        rec["meta"] = np.array([m, n, seed, aseed], dtype=np.int64)
        rec["meta_f"] = np.array([two_prob, wild], dtype=np.float64)
        rec["reward_kind"] = np.array(reward)
        rec["sched_coins"] = np.asarray(proxy.coins, np.float64)
        rec["sched_offsets"] = np.asarray(proxy.offsets, np.int64)
        rec["sched_perms"] = np.stack([first_perm] + proxy.perms)
        # only randfloat[0:16] is ever read (game_numba.py:207); keep the whole row anyway
        rec["sched_floats"] = np.stack([first_float] + proxy.floats)
        np.savez_compressed(os.path.join(OUT_DIR, f"schedule_{name}.npz"), **rec)
        print(f"schedule {name}: refreshes={len(proxy.perms)}")


def gen_gae() -> None:
    """compute_gae (gae.py:7-68) of the live reference on random fp32 inputs, with a stand-in critic that returns
    pre-drawn values (first call: v0, second call: v1)."""
    import torch
    from ml2048.gae import compute_gae
    from ml2048.stats import TensorStats

    torch.manual_seed(0)
    out = {}
    for tag, (u, s_, g), gamma, lam in (("a", (2, 16, 257), 0.997, 0.95), ("b", (1, 64, 33), 0.9, 0.5), ("c", (3, 5, 1024), 1.0, 1.0)):
        v0 = torch.randn(u, s_, g) * 50
        v1 = torch.randn(u, s_, g) * 50
        reward = (torch.randint(0, 64, (u, s_, g)) * 4).float()
        terminated = torch.rand(u, s_, g) < 0.05

        class Critic:
            def __init__(self):
                self.calls = 0

            def eval_value(self, state, valid):
                self.calls += 1
                return v0 if self.calls == 1 else v1

        data = {"state": torch.zeros(u, s_, g, 16, dtype=torch.int8), "valid_actions": torch.ones(u, s_, g, 4, dtype=torch.bool),
                "reward": reward, "next_state": torch.zeros(u, s_, g, 16, dtype=torch.int8),
                "next_valid_actions": torch.ones(u, s_, g, 4, dtype=torch.bool), "terminated": terminated,
                "adv": torch.zeros(u, s_, g)}
        compute_gae(Critic(), data, gamma=gamma, lambda_=lam, tensor_stats=TensorStats())
        for k, v in (("v0", v0), ("v1", v1), ("reward", reward), ("terminated", terminated), ("adv", data["adv"])):
            out[f"{tag}_{k}"] = v.numpy()
        out[f"{tag}_params"] = np.array([gamma, lam], np.float64)
    np.savez_compressed(os.path.join(OUT_DIR, "gae.npz"), **out)


def gen_runner() -> None:
    """The reference's OWN rollout stack -- VecRunner + RunnerStats + ReplayRecorder (runner.py, replay.py) -- driven by
    a deterministic policy, so that the device-side runner / statistics / episode log / trajectory capture can be held
    to the reference's code rather than to a restatement of it."""
    import torch
    from ml2048.policy import Policy
    from ml2048.replay import ReplayRecorder
    from ml2048.runner import RunnerStats, VecRunner

    class CyclingPolicy(Policy):
        """k-th valid action with k = (7 t + 13 slot) mod nvalid; action 0 when nothing is valid."""

        def __init__(self):
            super().__init__()
            self.t = 0

        def sample_actions(self, state, valid_actions, *, generator=None):
            m = valid_actions.shape[0]
            nvalid = valid_actions.sum(dim=1)
            k = (7 * self.t + 13 * torch.arange(m)) % nvalid.clamp(min=1)
            rank = torch.cumsum(valid_actions.long(), dim=1) - 1
            hit = valid_actions & (rank == k[:, None])
            actions = torch.where(nvalid > 0, hit.long().argmax(dim=1), torch.zeros(m, dtype=torch.long))
            self.t += 1
            return actions, torch.zeros(m)

    m, steps, seed = 384, 900, 11
    vg = ref.VecGame(m, ref.reward_fn_improved)
    vg.reset(seed)
    runner = VecRunner(vg, 16, sample_device="cpu")
    stats = RunnerStats()
    rec = ReplayRecorder(10**9, 10**9, segment_size=64)
    runner.add_callback(VecRunner.EVENT_PREPARED, rec.on_prepared)
    runner.add_callback(VecRunner.EVENT_STEPPED, rec.on_stepped)
    runner.add_callback(VecRunner.EVENT_STEPPED, stats.on_stepped)
    runner.step_many(CyclingPolicy(), steps)
    bufs = sorted(rec.ready_buffers, key=lambda b: b.id)
    out = {
        "meta": np.array([m, steps, seed], np.int64),
        "stats_counts": stats.counts.astype(np.int64),
        "stats_terminated": np.array(int(stats.terminated_count), np.int64),
        "summary_live": np.array([[a, b] for a, b, _ in vg.summary()], np.int64),
        "buf_id": np.array([b.id for b in bufs], np.int64),
        "buf_steps": np.array([b.steps for b in bufs], np.int64),
        "buf_maxcell": np.array([b.maxcell for b in bufs], np.int64),
        "buf_score": np.array([b.score for b in bufs], np.float32),
        "game_count": np.array(vg._game_count, np.int64),
    }
    keep = [b for b in bufs if b.id in (0, 1, 100, 383, 384, 500, 1000)]
    for b in keep:
        st, ac, sc = b.contiguous_result()
        out[f"traj_{b.id}_state"], out[f"traj_{b.id}_action"], out[f"traj_{b.id}_score"] = st, ac, sc
    out["traj_ids"] = np.array([b.id for b in keep], np.int64)
    np.savez_compressed(os.path.join(OUT_DIR, "runner_stack.npz"), **out)
    print(f"runner stack: {len(bufs)} finished recorded games, terminated_count={int(stats.terminated_count)}")


def gen_random_stats() -> None:
    """Episode statistics of the LIVE reference under the uniform-over-valid policy (policy/random.py:24 semantics), one run
    per seed: M = 8192 games x 1200 steps (the BASELINE.md section 2 protocol, whose seed-2024 figures are reproduced by the
    first entry).  All games of a run share the spawn tables -- in particular the 2-vs-4 choice is tied to the CELL for a
    whole table epoch (game_numba.py:207) -- so episodes within a run are not independent draws and the statistics vary
    between seeds by more than multinomial noise.  The Philox-mode tests measure that over-dispersion from these runs."""
    from .rollout import pick_actions

    seeds = [2024] + list(range(100, 115))
    m, steps = 8192, 1200
    hist = np.zeros((len(seeds), 20), np.int64)
    episodes = np.zeros(len(seeds), np.int64)
    step_sum = np.zeros(len(seeds), np.int64)
    score_sum = np.zeros(len(seeds), np.float64)
    for i, seed in enumerate(seeds):
        vg = ref.VecGame(m)
        vg.reset(seed)
        rng = np.random.default_rng(seed + 1)
        for _ in range(steps):
            vg.prepare()
            res = vg.step(pick_actions(vg.observations()[1], rng, 0.0))
            term = res["terminated"] != 0
            if term.any():
                np.add.at(hist[i], res["state"][term].max(axis=1), 1)
                episodes[i] += int(term.sum())
                step_sum[i] += int(res["step"][term].sum())
                score_sum[i] += float(res["score"][term].astype(np.float64).sum())
        print(f"random stats seed {seed}: episodes={episodes[i]} mean steps={step_sum[i]/episodes[i]:.2f} "
              f"mean score={score_sum[i]/episodes[i]:.1f} hist={hist[i][hist[i] > 0].tolist()}")
    np.savez_compressed(os.path.join(OUT_DIR, "random_policy_stats.npz"), seeds=np.array(seeds, np.int64),
                        meta=np.array([m, steps], np.int64), hist=hist, episodes=episodes, step_sum=step_sum, score_sum=score_sum)


def main() -> None:
    os.makedirs(OUT_DIR, exist_ok=True)
    print("numba", numba.__version__, "numpy", np.__version__)
    only = set(sys.argv[1:])
    for name, fn in (("kat", gen_kat), ("lines", gen_line_table), ("boards", gen_boards), ("rollouts", gen_rollouts),
                     ("schedules", gen_schedules), ("gae", gen_gae), ("runner", gen_runner),
                     ("random_stats", gen_random_stats)):
        if not only or name in only:
            fn()
    with open(os.path.join(OUT_DIR, "PROVENANCE.txt"), "w") as fh:
        fh.write("generated by oracle/gen_golden.py from the live reference at /root/reference/src\n")
        fh.write(f"numba {numba.__version__} numpy {np.__version__} python {sys.version.split()[0]}\n")


if __name__ == "__main__":
    main()
