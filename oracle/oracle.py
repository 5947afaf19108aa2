"""
oracle/oracle.py -- Python face of the CPU oracle (TEST INFRASTRUCTURE, NOT PRODUCT).

Loads ``oracle/vecgame_oracle.c`` (compiled by ``oracle/build_oracle.py``) through ctypes and
wraps it in ``OracleVecGame``, a class with the reference ``VecGame`` surface
(reference: src/ml2048/game_numba.py:522-698).  The host-side random schedule (numpy
``Generator(PCG64)`` draws) is restated here exactly in the reference's order:

    reset(seed)   default_rng(seed); randperm[:] = arange(16); permuted(out=); random(f32, out=)   :606-611, :589-591
    prepare()     random() >= 0.9 or rand_step >= 1024 -> rand_step = 0, table refresh; integers(0, 1024)   :622-626
    step()        integers(0, 1024); kernel seed = rand_step + offset; rand_step += 1                :670, :681, :685

Only tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference) may import
this module.  Parity status: PINNED against the live reference via tests/golden/ (see
oracle/gen_golden.py and tests/test_oracle_golden.py).
"""

from __future__ import annotations

import ctypes
import os
from typing import Any, Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "liboracle.so")

RAND_ROWS = 1024  # VecGame._RAND_SIZE, game_numba.py:533

REWARD_KINDS = {"normal": 0, "improved": 1, "rank": 2, "maxcell": 3}

# game_numba.py:537-550 -- same field order, align=True gives itemsize 64
DATA_DTYPE = np.dtype(
    [
        ("id", np.int32, ()),
        ("step", np.int32, ()),
        ("score", np.float32, ()),
        ("reward", np.float32, ()),
        ("board", np.uint8, (16,)),
        ("merged", np.uint8, (16,)),
        ("valid_actions", np.uint8, (4,)),
        ("terminated", np.uint8, ()),
        ("invalid", np.uint8, ()),
        ("_padding", np.uint8, 10),
    ],
    align=True,
)
assert DATA_DTYPE.itemsize == 64

_lib = None


def load_lib() -> ctypes.CDLL:
    """Load (building on first use if a compiler is present) the oracle shared library."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        from . import build_oracle

        build_oracle.build()
    lib = ctypes.CDLL(LIB_PATH)
    c = ctypes
    vp, i64, i32, dbl = c.c_void_p, c.c_int64, c.c_int, c.c_double
    lib.orc_vec_step.argtypes = [vp, i64, vp, vp, i32, dbl, i64, vp, vp, i64]
    lib.orc_vec_step.restype = None
    lib.orc_snapshot_prev.argtypes = [vp, i64, vp, vp]
    lib.orc_snapshot_prev.restype = None
    lib.orc_prepare.argtypes = [vp, i64, i64, dbl, vp, vp, i64, vp, vp]
    lib.orc_prepare.restype = i64
    lib.orc_reset.argtypes = [vp, i64]
    lib.orc_reset.restype = None
    lib.orc_onehot.argtypes = [vp, i64, i64, vp]
    lib.orc_onehot.restype = None
    lib.orc_terminated_hist.argtypes = [vp, i64, vp]
    lib.orc_terminated_hist.restype = i64
    lib.orc_live_hist.argtypes = [vp, i64, vp]
    lib.orc_live_hist.restype = None
    lib.orc_board_move.argtypes = [vp, vp, i64]
    lib.orc_board_move.restype = None
    lib.orc_board_valid.argtypes = [vp, vp]
    lib.orc_board_valid.restype = i32
    lib.orc_board_reward.argtypes = [i32, vp, vp, vp]
    lib.orc_board_reward.restype = dbl
    lib.orc_line_push.argtypes = [vp, i32, vp]
    lib.orc_line_push.restype = None
    lib.orc_line_flags.argtypes = [i32, i32, i32, i32, vp, vp]
    lib.orc_line_flags.restype = None
    lib.orc_board_spawn.argtypes = [vp, vp, i64, i64, vp, dbl, i32]
    lib.orc_board_spawn.restype = i32
    lib.orc_random_valid_actions.argtypes = [vp, i64, c.c_uint64, vp]
    lib.orc_random_valid_actions.restype = None
    lib.orc_num_threads.argtypes = []
    lib.orc_num_threads.restype = i32
    lib.orc_set_num_threads.argtypes = [i32]
    lib.orc_set_num_threads.restype = None
    _lib = lib
    return lib


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


# --------------------------------------------------------------------------------------------
# single-board helpers (known-answer and exhaustive tests)
# --------------------------------------------------------------------------------------------


def board_move(board: np.ndarray, action: int) -> tuple[np.ndarray, np.ndarray]:
    """Return (new_board, merged) for one 16-cell board; follows game_numba.py:125-134."""
    lib = load_lib()
    b = np.ascontiguousarray(board, dtype=np.uint8).copy()
    m = np.zeros(16, dtype=np.uint8)
    lib.orc_board_move(_ptr(b), _ptr(m), int(action))
    return b, m


def board_valid(board: np.ndarray) -> np.ndarray:
    """Valid-action mask (L, R, U, D) as uint8[4]; follows game_numba.py:259-289."""
    lib = load_lib()
    b = np.ascontiguousarray(board, dtype=np.uint8)
    out = np.zeros(4, dtype=np.uint8)
    lib.orc_board_valid(_ptr(b), _ptr(out))
    return out


def board_reward(kind: str, state: np.ndarray, prev: np.ndarray, merged: np.ndarray) -> float:
    """One of the four reward functions, game_numba.py:408-504."""
    lib = load_lib()
    s = np.ascontiguousarray(state, dtype=np.uint8)
    p = np.ascontiguousarray(prev, dtype=np.uint8)
    m = np.ascontiguousarray(merged, dtype=np.uint8)
    return float(lib.orc_board_reward(REWARD_KINDS[kind], _ptr(s), _ptr(p), _ptr(m)))


def line_push(line: np.ndarray, toward_last: bool) -> tuple[np.ndarray, np.ndarray]:
    """Push one 4-cell line (toward cell 0, or toward cell 3); returns (line, buckets[18])."""
    lib = load_lib()
    l = np.ascontiguousarray(line, dtype=np.uint8).copy()
    buckets = np.zeros(18, dtype=np.uint8)
    lib.orc_line_push(_ptr(l), int(bool(toward_last)), _ptr(buckets))
    return l, buckets


def line_flags(n1: int, n2: int, n3: int, n4: int) -> tuple[bool, bool]:
    """(movable toward position 1, movable toward position 4); game_numba.py:215-256."""
    lib = load_lib()
    f = ctypes.c_int(0)
    b = ctypes.c_int(0)
    lib.orc_line_flags(n1, n2, n3, n4, ctypes.byref(f), ctypes.byref(b))
    return bool(f.value), bool(b.value)


def onehot(boards: np.ndarray) -> np.ndarray:
    """(M,16) u8 -> (M,16,16) f32, class-major; policy/_network.py:86-95."""
    lib = load_lib()
    b = np.ascontiguousarray(boards, dtype=np.uint8)
    out = np.empty((b.shape[0], 16, 16), dtype=np.float32)
    lib.orc_onehot(_ptr(b), 16, b.shape[0], _ptr(out))
    return out


# --------------------------------------------------------------------------------------------
# host random schedule
# --------------------------------------------------------------------------------------------


class NumpySchedule:
    """The reference's own host draws, restated (game_numba.py:589-591, 606-611, 622-626, 670)."""

    def __init__(self, seed: Optional[int]):
        self._rand = np.random.default_rng(seed)

    def refresh_tables(self, randperm: np.ndarray, randfloat: np.ndarray) -> None:
        self._rand.permuted(randperm, axis=1, out=randperm)
        self._rand.random(dtype=randfloat.dtype, out=randfloat)

    def refresh_coin(self) -> float:
        return self._rand.random()

    def offset(self) -> int:
        return int(self._rand.integers(0, RAND_ROWS))


class RecordedSchedule:
    """Replays draws recorded from a live reference instance (see oracle/gen_golden.py).

    ``coins``  float64[n_prepare]        the random() of every prepare()
    ``offsets`` int64[n_prepare+n_step]   every integers(0,1024), in call order
    ``tables``  list of (randperm u8 (1024,16), randfloat f32 (1024,)) in refresh order,
                the first entry being the tables right after reset().
    """

    def __init__(self, coins: np.ndarray, offsets: np.ndarray, perms: np.ndarray, floats: np.ndarray):
        self._coins = list(np.asarray(coins, dtype=np.float64))
        self._offsets = list(np.asarray(offsets, dtype=np.int64))
        self._perms = list(perms)
        self._floats = list(floats)

    def refresh_tables(self, randperm: np.ndarray, randfloat: np.ndarray) -> None:
        randperm[...] = self._perms.pop(0)
        randfloat[...] = self._floats.pop(0)

    def refresh_coin(self) -> float:
        return float(self._coins.pop(0))

    def offset(self) -> int:
        return int(self._offsets.pop(0))


# --------------------------------------------------------------------------------------------
# the environment
# --------------------------------------------------------------------------------------------


class OracleVecGame:
    """CPU oracle with the reference ``VecGame`` surface (game_numba.py:522-698)."""

    _RAND_SIZE = RAND_ROWS
    _DATA_DTYPE = DATA_DTYPE

    def __init__(self, size: int, reward_fn: Any = None, *, two_prob: float = 0.8, reuse_state: bool = False):
        if size <= 0:
            raise ValueError(f"size={size}")  # game_numba.py:561-562
        self._lib = load_lib()
        self._size = int(size)
        self._two_prob = float(two_prob)
        self._reuse_state = reuse_state
        self._reward_kind = _reward_kind(reward_fn)
        self._data = np.empty((size,), dtype=DATA_DTYPE)
        self._prev_state = np.empty((size, 16), dtype=np.uint8)
        self._prev_valid_actions = np.empty((size, 4), dtype=np.uint8)
        self._randperm = np.empty((RAND_ROWS, 16), dtype=np.uint8)
        self._randfloat = np.empty((RAND_ROWS,), dtype=np.float32)
        self._rand_step = 0
        self._game_count = 0  # survives reset(), game_numba.py:582
        self._indices = np.empty((size,), dtype=np.int64)
        self._schedule: Any = None
        self.reset()

    # -- reference surface ------------------------------------------------------------------

    def reset(self, seed: Optional[int] = None, *, schedule: Any = None) -> None:
        self._schedule = schedule if schedule is not None else NumpySchedule(seed)
        self._rand_step = 0
        self._randperm[:, :] = np.arange(16).reshape((1, 16))
        self._schedule.refresh_tables(self._randperm, self._randfloat)
        self._lib.orc_reset(_ptr(self._data), self._size)
        self._prev_state.fill(0)
        self._prev_valid_actions.fill(0)

    def observations(self) -> tuple[np.ndarray, np.ndarray]:
        return self._data["board"], self._data["valid_actions"]

    def prepare(self) -> tuple[np.ndarray]:
        if self._schedule.refresh_coin() >= 0.9 or self._rand_step >= RAND_ROWS:
            self._rand_step = 0
            self._schedule.refresh_tables(self._randperm, self._randfloat)
        rand_offset = self._schedule.offset()
        gc = ctypes.c_int64(self._game_count)
        n = self._lib.orc_prepare(
            _ptr(self._data),
            self._size,
            self._rand_step + rand_offset,
            self._two_prob,
            _ptr(self._randperm),
            _ptr(self._randfloat),
            RAND_ROWS,
            ctypes.byref(gc),
            _ptr(self._indices),
        )
        self._game_count = gc.value
        return (self._indices[:n].copy(),)

    def step(self, actions: np.ndarray) -> dict[str, np.ndarray]:
        assert actions.shape == (self._size,), actions.shape  # game_numba.py:668
        acts = np.ascontiguousarray(actions, dtype=np.int64)
        rand_offset = self._schedule.offset()
        self._lib.orc_snapshot_prev(_ptr(self._data), self._size, _ptr(self._prev_state), _ptr(self._prev_valid_actions))
        self._lib.orc_vec_step(
            _ptr(self._data),
            self._size,
            _ptr(acts),
            _ptr(self._prev_state),
            self._reward_kind,
            self._two_prob,
            self._rand_step + rand_offset,
            _ptr(self._randperm),
            _ptr(self._randfloat),
            RAND_ROWS,
        )
        self._rand_step += 1
        d = self._data
        return {
            "state": d["board"],
            "valid_actions": d["valid_actions"],
            "merged": d["merged"],
            "step": d["step"],
            "reward": d["reward"],
            "score": d["score"],
            "terminated": d["terminated"],
            "invalid": d["invalid"],
            "prev_state": self._prev_state,
            "prev_valid_actions": self._prev_valid_actions,
        }

    def summary(self) -> list[Any]:
        counts = np.zeros(20, dtype=np.int64)
        self._lib.orc_live_hist(_ptr(self._data), self._size, _ptr(counts))
        total = int(counts.sum())
        entries = [(2 ** int(k), int(counts[k]), counts[k] / total) for k in range(20) if counts[k]]
        entries.sort(key=lambda s: s[0], reverse=True)
        return entries

    def random_valid_actions(self, seed: int, out: Optional[np.ndarray] = None) -> np.ndarray:
        """Uniform-over-valid actions for every game (benchmark policy), computed in C."""
        if out is None:
            out = np.empty((self._size,), dtype=np.int64)
        self._lib.orc_random_valid_actions(_ptr(self._data), self._size, seed & 0xFFFFFFFFFFFFFFFF, _ptr(out))
        return out

    # -- adjacent statistic (RunnerStats, runner.py:120-166) ----------------------------------

    def terminated_hist(self) -> tuple[np.ndarray, int]:
        counts = np.zeros(20, dtype=np.int64)
        n = self._lib.orc_terminated_hist(_ptr(self._data), self._size, _ptr(counts))
        return counts, int(n)


def _reward_kind(reward_fn: Any) -> int:
    if reward_fn is None:
        return 0
    if isinstance(reward_fn, str):
        name = reward_fn
    else:
        name = getattr(reward_fn, "__name__", None) or getattr(getattr(reward_fn, "py_func", None), "__name__", "")
    name = name.replace("reward_fn_", "")
    if name not in REWARD_KINDS:
        raise ValueError(f"unknown reward_fn {reward_fn!r}")
    return REWARD_KINDS[name]


def random_valid_actions(valid: np.ndarray, u: np.ndarray) -> np.ndarray:
    """Uniform choice among valid actions (semantics of policy/random.py:24) from uniforms ``u``:
    the floor(u * nvalid)-th valid action, action 0 when none is valid."""
    v = valid.astype(bool)
    nvalid = v.sum(axis=1)
    k = np.minimum((u * nvalid).astype(np.int64), np.maximum(nvalid - 1, 0))
    rank = np.cumsum(v, axis=1) - 1
    hit = v & (rank == k[:, None])
    return np.where(nvalid > 0, hit.argmax(axis=1), 0).astype(np.int64)
