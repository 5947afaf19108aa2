"""Host-side random schedule of the environment.

The reference draws a few state-independent numbers from a numpy ``Generator(PCG64)`` on the host in
every ``prepare()``/``step()`` (reference: src/ml2048/game_numba.py:589-591, 606-611, 622-626, 670).
``NumpySchedule`` performs exactly those draws in exactly that order, so that with the same seed the
tables and offsets -- and hence every board -- equal the reference's.  ``RecordedSchedule`` consumes
draws recorded from a live reference instance instead (the "replay of pre-drawn uniforms" mode: it
does not depend on numpy's stream at all).
"""

from __future__ import annotations

from typing import Optional, Sequence

import numpy as np

RAND_ROWS = 1024  # VecGame._RAND_SIZE, game_numba.py:533


class NumpySchedule:
    def __init__(self, seed: Optional[int]):
        self._rand = np.random.default_rng(seed)  # game_numba.py:607

    def refresh_tables(self, randperm: np.ndarray, randfloat: np.ndarray) -> None:
        # game_numba.py:589-591 (_reset_rand): in-place, cumulative permutation of every row
        self._rand.permuted(randperm, axis=1, out=randperm)
        self._rand.random(dtype=randfloat.dtype, out=randfloat)

    def refresh_coin(self) -> float:
        return self._rand.random()  # game_numba.py:622

    def offset(self) -> int:
        return int(self._rand.integers(0, RAND_ROWS))  # game_numba.py:626 and :670


class RecordedSchedule:
    """Draws recorded from the reference: ``coins`` (one per prepare), ``offsets`` (one per prepare and
    one per step, in call order), ``perms``/``floats`` (tables after reset() and after every refresh)."""

    def __init__(self, coins: Sequence[float], offsets: Sequence[int], perms: Sequence[np.ndarray], floats: Sequence[np.ndarray]):
        self._coins = [float(x) for x in coins]
        self._offsets = [int(x) for x in offsets]
        self._perms = [np.asarray(p, dtype=np.uint8) for p in perms]
        self._floats = [np.asarray(f, dtype=np.float32) for f in floats]

    def refresh_tables(self, randperm: np.ndarray, randfloat: np.ndarray) -> None:
        if not self._perms:
            raise RuntimeError("recorded schedule exhausted (tables)")
        randperm[...] = self._perms.pop(0)
        randfloat[...] = self._floats.pop(0)

    def refresh_coin(self) -> float:
        if not self._coins:
            raise RuntimeError("recorded schedule exhausted (coins)")
        return self._coins.pop(0)

    def offset(self) -> int:
        if not self._offsets:
            raise RuntimeError("recorded schedule exhausted (offsets)")
        return self._offsets.pop(0)
