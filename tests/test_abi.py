"""CPU-side checks of the drop-in boundary: the C-ABI library loads without a GPU and exports every
symbol include/ml2048_b200.h declares; the ctypes mirrors match the C structs; host-side logic."""

from __future__ import annotations

import ctypes
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "ml2048_b200.h")


@pytest.fixture(scope="module")
def lib():
    from ml2048_b200 import _lib, build

    build.build()  # nvcc cross-compiles without a GPU; no-op when up to date
    return _lib.load()


def test_every_declared_symbol_is_exported(lib):
    from ml2048_b200 import _lib

    text = open(HEADER).read()
    declared = set(re.findall(r"ML2048_API\s+[\w\s\*]+?\b(ml2048_\w+)\s*\(", text))
    assert len(declared) >= 13
    assert declared == set(_lib.SYMBOLS), declared ^ set(_lib.SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.ml2048_abi_version() == _lib.ABI_VERSION


def test_struct_layouts_match_header(lib, tmp_path):
    from ml2048_b200 import _lib

    src = tmp_path / "sz.c"
    src.write_text(
        f'#include "{HEADER}"\n#include <stdio.h>\n#include <stddef.h>\n'
        "int main(void){printf(\"%zu %zu %zu %zu %zu %zu\\n\", sizeof(ml2048_step_args), sizeof(ml2048_prepare_args),"
        " sizeof(ml2048_stats), offsetof(ml2048_step_args, randperm_keys), offsetof(ml2048_step_args, stats),"
        " offsetof(ml2048_prepare_args, scratch));return 0;}\n"
    )
    exe = tmp_path / "sz"
    subprocess.run(["gcc", str(src), "-o", str(exe)], check=True)
    got = [int(x) for x in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert got == [
        ctypes.sizeof(_lib.StepArgs),
        ctypes.sizeof(_lib.PrepareArgs),
        _lib.STATS_WORDS * 8,
        _lib.StepArgs.randperm_keys.offset,
        _lib.StepArgs.stats.offset,
        _lib.PrepareArgs.scratch.offset,
    ]


def test_argument_errors_without_gpu(lib):
    from ml2048_b200 import _lib

    assert lib.ml2048_step(None, None) == -1
    a = _lib.StepArgs()
    assert lib.ml2048_step(ctypes.byref(a), None) == -5  # struct_size not set
    a.struct_size = ctypes.sizeof(_lib.StepArgs)
    assert lib.ml2048_step(ctypes.byref(a), None) == -3  # num_games == 0
    a.num_games = 8
    assert lib.ml2048_step(ctypes.byref(a), None) == -1  # null boards
    # step and score must be the halves of one array of 8-byte {step, score} records (checked before anything is launched)
    a.board_in, a.board_out, a.valid_out = 0x10000, 0x20000, 0x30000
    a.reward, a.terminated, a.invalid = 0x40000, 0x50000, 0x60000
    a.step, a.score = 0x70000, 0x80000
    assert lib.ml2048_step(ctypes.byref(a), None) == -2
    a.step, a.score = 0x70004, 0x70008
    assert lib.ml2048_step(ctypes.byref(a), None) == -2  # records are 8-byte aligned
    assert lib.ml2048_prepare_scratch_ints(0) == 0
    assert lib.ml2048_prepare_scratch_ints(4096) == 4
    assert lib.ml2048_prepare_scratch_ints(4097) == 4


def test_two_mask_uses_double_compare(lib):
    # game_numba.py:207: float32(0.8) = 0.800000011920929 is NOT < 0.8 as a double: that cell spawns a 4
    rf = np.zeros(16, np.float32)
    rf[3] = np.float32(0.8)
    rf[5] = np.nextafter(np.float32(0.8), np.float32(0))
    rf[7] = 0.9
    mask = lib.ml2048_two_mask(rf.ctypes.data, 0.8)
    assert mask == (0xFFFF & ~(1 << 3) & ~(1 << 7))
    assert lib.ml2048_two_threshold(0.0) == 0
    assert lib.ml2048_two_threshold(1.0) == 0xFFFFFFFF
    assert lib.ml2048_two_threshold(0.5) == 0x80000000


def test_host_philox_known_answers_and_epoch_draws(lib):
    """Host Philox of the library (no GPU): Random123 known answers, and the Philox-mode table-epoch draws."""
    out4 = np.zeros(4, np.uint32)
    for ctr, key, want in (((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
                           ((0xFFFFFFFF,) * 4, (0xFFFFFFFF,) * 2, (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
                           ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0),
                            (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1))):
        c, k = np.array(ctr, np.uint32), np.array(key, np.uint32)
        lib.ml2048_philox4x32_10(c.ctypes.data, k.ctypes.data, out4.ctypes.data)
        assert tuple(int(x) for x in out4) == want
    out2 = np.zeros(2, np.uint32)
    for ctr, key, want in (((0, 0), 0, (0xFF1DAE59, 0x6CD10DF2)), ((0xFFFFFFFF,) * 2, 0xFFFFFFFF, (0x2C3F628B, 0xAB4FD7AD)),
                           ((0x243F6A88, 0x85A308D3), 0x13198A2E, (0xDD7CE038, 0xF62A4C12))):
        c = np.array(ctr, np.uint32)
        lib.ml2048_philox2x32_10(c.ctypes.data, key, out2.ctypes.data)
        assert tuple(int(x) for x in out2) == want
    coin, mask = ctypes.c_uint32(), ctypes.c_uint32()
    ones = refresh = 0
    n = 4000
    for counter in range(n):
        lib.ml2048_philox_epoch_draws(99, counter, 0.8, ctypes.byref(coin), ctypes.byref(mask))
        assert mask.value < (1 << 16)
        ones += bin(mask.value).count("1")
        refresh += coin.value >= int(0.9 * 2**32)
    assert abs(ones / (16 * n) - 0.8) < 0.01 and abs(refresh / n - 0.1) < 0.02
    lib.ml2048_philox_epoch_draws(99, 7, 0.8, ctypes.byref(coin), ctypes.byref(mask))
    first = (coin.value, mask.value)
    lib.ml2048_philox_epoch_draws(99, 7, 0.8, ctypes.byref(coin), ctypes.byref(mask))
    assert (coin.value, mask.value) == first  # a pure function of (seed, counter): every shard draws the same epoch
    lib.ml2048_philox_epoch_draws(99, 7, 1.0, ctypes.byref(coin), ctypes.byref(mask))
    assert mask.value == 0xFFFF
    lib.ml2048_philox_epoch_draws(99, 7, 0.0, ctypes.byref(coin), ctypes.byref(mask))
    assert mask.value == 0


def test_reward_selection_by_identity():
    import ml2048_b200
    from ml2048_b200.rewards import reward_kind

    assert reward_kind(None) == 0
    assert reward_kind(ml2048_b200.reward_fn_improved) == 1
    assert reward_kind("rank") == 2

    def reward_fn_maxcell(state, prev, merged):  # the reference's own function object is accepted by name
        return 0.0

    assert reward_kind(reward_fn_maxcell) == 3
    with pytest.raises(ValueError):
        reward_kind(lambda s, p, m: 0.0)


def test_no_cpu_fallback():
    import torch

    import ml2048_b200

    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ml2048_b200.VecGame(8)


def test_host_schedule_matches_oracle_schedule():
    """The product's host draws (host_rng.py) and the oracle's are the same numpy calls in the same order."""
    from ml2048_b200.host_rng import NumpySchedule
    from oracle.oracle import NumpySchedule as OracleSchedule

    for seed in (0, 1, 123):
        a, b = NumpySchedule(seed), OracleSchedule(seed)
        pa = np.tile(np.arange(16, dtype=np.uint8), (1024, 1))
        pb = pa.copy()
        fa, fb = np.empty(1024, np.float32), np.empty(1024, np.float32)
        for _ in range(3):
            a.refresh_tables(pa, fa)
            b.refresh_tables(pb, fb)
            assert np.array_equal(pa, pb) and np.array_equal(fa, fb)
            assert a.refresh_coin() == b.refresh_coin()
            assert a.offset() == b.offset()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "ml2048_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert not re.search(r"(import\s+oracle|from\s+oracle|from\s+\.+oracle|liboracle|oracle/)", text), f"{f} reaches into oracle/"
    code = "import sys; import ml2048_b200, ml2048_b200.vecgame, ml2048_b200.sharding; assert not any(m.startswith('oracle') for m in sys.modules)"
    subprocess.run([sys.executable, "-c", code], check=True, cwd=ROOT)


def test_pack_randperm_keys(lib):
    rng = np.random.default_rng(0)
    perm = np.argsort(rng.random((1024, 16)), axis=1).astype(np.uint8)
    keys = np.zeros_like(perm)
    assert lib.ml2048_pack_randperm_keys(perm.ctypes.data, keys.ctypes.data, 1024) == 0
    want = np.empty_like(perm)
    want[np.arange(1024)[:, None], perm] = (np.arange(16, dtype=np.uint8) * 16)[None, :] + perm
    np.testing.assert_array_equal(keys, want)
    assert (keys & 15 == np.arange(16)).all()
    bad = perm.copy()
    bad[3, 0] = bad[3, 1]
    assert lib.ml2048_pack_randperm_keys(bad.ctypes.data, keys.ctypes.data, 1024) == -4
    assert lib.ml2048_pack_randperm_keys(None, keys.ctypes.data, 1024) == -1


def test_unpack_flags_host(lib):
    """ml2048_unpack_flags (host code): one packed byte per game -> valid_actions[4], terminated, invalid; every size class of the
    vector / word / scalar loops, any subset of the outputs, any thread count; bytes next to the outputs stay untouched."""
    rng = np.random.default_rng(1)
    for n in (1, 7, 8, 9, 63, 64, 65, 1000, (1 << 17) + 5, (1 << 20) + 3):
        p = rng.integers(0, 256, n).astype(np.uint8)
        for threads in (0, 1, 3, 16):
            v = np.full((n + 2, 4), 7, np.uint8)
            t = np.full(n + 2, 7, np.uint8)
            i = np.full(n + 2, 7, np.uint8)
            lib.ml2048_unpack_flags(p.ctypes.data, n, v[1:].ctypes.data, t[1:].ctypes.data, i[1:].ctypes.data, threads)
            np.testing.assert_array_equal(v[1:-1], (p[:, None] >> np.arange(4)) & 1)
            np.testing.assert_array_equal(t[1:-1], (p >> 4) & 1)
            np.testing.assert_array_equal(i[1:-1], (p >> 5) & 1)
            assert (v[0] == 7).all() and (v[-1] == 7).all() and t[0] == t[-1] == i[0] == i[-1] == 7
        t = np.full(n, 7, np.uint8)
        lib.ml2048_unpack_flags(p.ctypes.data, n, None, t.ctypes.data, None, 2)
        np.testing.assert_array_equal(t, (p >> 4) & 1)
    lib.ml2048_unpack_flags(None, 10, None, None, None, 1)  # nothing to do, no crash
    assert lib.ml2048_pack_flags(None, None, None, None, 10, None) == -1
    assert lib.ml2048_pack_flags(1 << 20, None, None, 1 << 21, 0, None) == -3
