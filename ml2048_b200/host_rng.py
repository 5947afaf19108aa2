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


class Pcg64Schedule:
    """The same draws as ``NumpySchedule`` from the same PCG64 stream, made by the C functions of the library
    (``ml2048_pcg64_*``, include/ml2048_b200.h) instead of numpy calls: numpy's ``Generator.permuted`` over the
    (1024, 16) table costs ~470 us per refresh, the C restatement ~170 us, and the scalar draws drop from 1-3 us to
    ~0.3 us each.  The generator is SEEDED by numpy (``default_rng(seed).bit_generator.state``), so seeds mean the
    same thing as in the reference."""

    def __init__(self, seed: Optional[int]):
        import ctypes as C

        from . import _lib

        self._C = C
        self._lib = _lib.load()
        st = np.random.default_rng(seed).bit_generator.state
        if st["bit_generator"] != "PCG64":
            raise RuntimeError("numpy's default bit generator is not PCG64")
        s, i = st["state"]["state"], st["state"]["inc"]
        m = (1 << 64) - 1
        self._g = _lib.Pcg64(s >> 64, s & m, i >> 64, i & m, int(st["has_uint32"]), int(st["uinteger"]))
        self._ref = C.byref(self._g)

    def refresh_tables(self, randperm: np.ndarray, randfloat: np.ndarray) -> None:
        assert randperm.dtype == np.uint8 and randperm.flags.c_contiguous and randfloat.dtype == np.float32
        self._lib.ml2048_pcg64_permuted_rows_u8(self._ref, randperm.ctypes.data, randperm.shape[0], randperm.shape[1])
        self._lib.ml2048_pcg64_random_f32(self._ref, randfloat.ctypes.data, randfloat.shape[0])

    def refresh_coin(self) -> float:
        return self._lib.ml2048_pcg64_random(self._ref)

    def offset(self) -> int:
        return self._lib.ml2048_pcg64_integers(self._ref, RAND_ROWS)

    def __deepcopy__(self, memo):
        other = Pcg64Schedule.__new__(Pcg64Schedule)
        other._C, other._lib = self._C, self._lib
        other._g = type(self._g).from_buffer_copy(self._g)
        other._ref = self._C.byref(other._g)
        return other


_fast_ok: Optional[bool] = None


def _fast_schedule_matches_numpy() -> bool:
    """One-time self check: the C generator must reproduce the installed numpy's draws exactly."""
    global _fast_ok
    if _fast_ok is None:
        try:
            a, b = NumpySchedule(20481), Pcg64Schedule(20481)
            pa = np.tile(np.arange(16, dtype=np.uint8), (RAND_ROWS, 1))
            pb = pa.copy()
            fa, fb = np.empty(RAND_ROWS, np.float32), np.empty(RAND_ROWS, np.float32)
            ok = True
            for _ in range(3):
                a.refresh_tables(pa, fa)
                b.refresh_tables(pb, fb)
                ok = ok and np.array_equal(pa, pb) and np.array_equal(fa, fb)
                for _ in range(5):
                    ok = ok and a.refresh_coin() == b.refresh_coin() and a.offset() == b.offset() and a.offset() == b.offset()
            _fast_ok = bool(ok)
        except Exception:  # noqa: BLE001 - any problem means: use numpy
            _fast_ok = False
    return _fast_ok


def make_schedule(seed: Optional[int]):
    """The host schedule for ``VecGame.reset(seed)``: the C generator when it provably equals numpy's, numpy otherwise."""
    if seed is not None and _fast_schedule_matches_numpy():
        return Pcg64Schedule(seed)
    return NumpySchedule(seed)


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
