"""The C restatement of numpy's PCG64 draws (ml2048_pcg64_*) against the installed numpy, long streams, CPU only."""

from __future__ import annotations

import copy

import numpy as np
import pytest


@pytest.mark.parametrize("seed", [0, 1, 123, 2024, 2**40 + 7])
def test_c_generator_reproduces_numpy_stream(seed):
    from ml2048_b200.host_rng import NumpySchedule, Pcg64Schedule

    a, b = NumpySchedule(seed), Pcg64Schedule(seed)
    pa = np.tile(np.arange(16, dtype=np.uint8), (1024, 1))
    pb = pa.copy()
    fa, fb = np.empty(1024, np.float32), np.empty(1024, np.float32)
    rng = np.random.default_rng(seed + 1)
    for it in range(400):  # interleave the draws in a random order, as prepare()/step() would
        k = rng.integers(0, 4)
        if k == 0 and it % 7 == 0:
            a.refresh_tables(pa, fa)
            b.refresh_tables(pb, fb)
            np.testing.assert_array_equal(pa, pb)
            np.testing.assert_array_equal(fa.view(np.uint32), fb.view(np.uint32))
        elif k == 1:
            assert a.refresh_coin() == b.refresh_coin()
        else:
            assert a.offset() == b.offset()
    c = copy.deepcopy(b)
    assert [c.offset() for _ in range(5)] == [b.offset() for _ in range(5)]


def test_make_schedule_prefers_the_verified_fast_path():
    from ml2048_b200 import host_rng

    assert host_rng._fast_schedule_matches_numpy() is True
    assert isinstance(host_rng.make_schedule(3), host_rng.Pcg64Schedule)
    assert isinstance(host_rng.make_schedule(None), host_rng.NumpySchedule)  # OS entropy: nothing to reproduce
