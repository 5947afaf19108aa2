"""
oracle/rollout.py -- lock-step rollout recorder shared by the golden generator and the tests.
TEST INFRASTRUCTURE, NOT PRODUCT.

Drives any object with the reference ``VecGame`` surface (the live reference, the C oracle or the
CUDA environment in its NumPy-compatible mode) through

    prepare() -> observations() -> pick actions -> step(actions)

exactly as ``VecRunner.step_once`` does (reference: src/ml2048/runner.py:74-109), and records every
``VecStepResult`` field (game_numba.py:507-519) plus the game ids and the reset indices, both in
full and as per-step CRC32 digests.
"""

from __future__ import annotations

import zlib
from typing import Any

import numpy as np

FIELDS = (
    # name, dtype, trailing shape
    ("state", np.uint8, (16,)),
    ("valid_actions", np.uint8, (4,)),
    ("merged", np.uint8, (16,)),
    ("step", np.int32, ()),
    ("reward", np.float32, ()),
    ("score", np.float32, ()),
    ("terminated", np.uint8, ()),
    ("invalid", np.uint8, ()),
    ("prev_state", np.uint8, (16,)),
    ("prev_valid_actions", np.uint8, (4,)),
)
DIGEST_NAMES = tuple(f[0] for f in FIELDS) + ("id", "reset_indices", "actions")


def _canon(a: Any, dtype: Any) -> np.ndarray:
    return np.ascontiguousarray(np.asarray(a), dtype=dtype)


def pick_actions(valid: np.ndarray, rng: np.random.Generator, wild: float) -> np.ndarray:
    """Uniform over valid actions (policy/random.py:24 semantics); a fraction ``wild`` of the games
    gets a uniformly random direction instead, valid or not, to exercise the invalid-move path."""
    v = np.asarray(valid).astype(bool)
    m = v.shape[0]
    u = rng.random(m)
    nvalid = v.sum(axis=1)
    k = np.minimum((u * nvalid).astype(np.int64), np.maximum(nvalid - 1, 0))
    rank = np.cumsum(v, axis=1) - 1
    hit = v & (rank == k[:, None])
    acts = np.where(nvalid > 0, hit.argmax(axis=1), 0).astype(np.int64)
    if wild > 0:
        sel = rng.random(m) < wild
        rnd = rng.integers(0, 4, size=m)
        acts = np.where(sel, rnd, acts).astype(np.int64)
    return acts


def get_ids(env: Any) -> np.ndarray:
    data = env._data
    return _canon(data["id"], np.int32)


def record_rollout(env: Any, steps: int, *, action_seed: int = 7, wild: float = 0.05, full: bool = True,
                   actions: np.ndarray | None = None) -> dict[str, np.ndarray]:
    """Run ``steps`` runner steps on ``env`` (already reset) and record everything.

    If ``actions`` (steps, M) is given it is replayed instead of sampling (used to drive a second
    implementation with the actions recorded from the first)."""
    m = env._size
    rng = np.random.default_rng(action_seed)
    out: dict[str, Any] = {name: [] for name, _, _ in FIELDS}
    out["id"] = []
    out["actions"] = []
    n_reset = []
    reset_idx = []
    digests = np.zeros((steps, len(DIGEST_NAMES)), dtype=np.uint32)

    for t in range(steps):
        (idx,) = env.prepare()
        idx = _canon(idx, np.int64)
        ids = get_ids(env).copy()
        _, valid = env.observations()
        if actions is None:
            acts = pick_actions(_canon(valid, np.uint8), rng, wild)
        else:
            acts = _canon(actions[t], np.int64)
        res = env.step(acts)
        row = {}
        for name, dt, _ in FIELDS:
            row[name] = _canon(res[name], dt).copy()
        row["id"] = ids
        row["reset_indices"] = idx
        row["actions"] = acts.astype(np.int8)
        for j, name in enumerate(DIGEST_NAMES):
            digests[t, j] = zlib.crc32(row[name].tobytes())
        n_reset.append(idx.size)
        reset_idx.append(idx.astype(np.int32))
        out["actions"].append(row["actions"])
        if full:
            for name, _, _ in FIELDS:
                out[name].append(row[name])
            out["id"].append(ids)

    rec: dict[str, np.ndarray] = {
        "digests": digests,
        "n_reset": np.asarray(n_reset, dtype=np.int64),
        "reset_indices": np.concatenate(reset_idx) if reset_idx else np.zeros(0, np.int32),
        "actions": np.stack(out["actions"]),
    }
    if full:
        for name, _, _ in FIELDS:
            rec[name] = np.stack(out[name])
        rec["id"] = np.stack(out["id"])
    # final snapshot, always kept
    board, valid = env.observations()
    rec["final_state"] = _canon(board, np.uint8).copy()
    rec["final_valid_actions"] = _canon(valid, np.uint8).copy()
    rec["final_id"] = get_ids(env).copy()
    rec["final_game_count"] = np.asarray(env._game_count, dtype=np.int64)
    return rec


def compare_rollouts(got: dict[str, np.ndarray], want: dict[str, np.ndarray]) -> None:
    """Assert bit-equality; reports the first step and field that differ."""
    np.testing.assert_array_equal(got["actions"], want["actions"], err_msg="actions")
    for name in ("n_reset", "reset_indices", "final_state", "final_valid_actions", "final_id", "final_game_count"):
        np.testing.assert_array_equal(got[name], want[name], err_msg=name)
    for name in [f[0] for f in FIELDS] + ["id"]:
        if name in got and name in want:
            g, w = got[name], want[name]
            if g.dtype.kind == "f":
                g, w = g.view(np.uint32), w.view(np.uint32)  # bit-exact, not approx
            if not np.array_equal(g, w):
                t = int(np.argwhere(g.reshape(g.shape[0], -1) != w.reshape(w.shape[0], -1))[0][0])
                bad = np.argwhere(g[t] != w[t])[:5]
                raise AssertionError(f"field {name!r} differs first at step {t}, positions {bad.tolist()}: "
                                     f"got {got[name][t][tuple(bad[0])]} want {want[name][t][tuple(bad[0])]}")
    gd, wd = got["digests"], want["digests"]
    if not np.array_equal(gd, wd):
        t, j = np.argwhere(gd != wd)[0]
        raise AssertionError(f"digest of {DIGEST_NAMES[j]!r} differs first at step {t}")
