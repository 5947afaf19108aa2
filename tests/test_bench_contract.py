"""bench.py's CPU-runnable arm must print exactly one JSON line with the contract's keys (the GPU arm is exercised on the
B200 box; its keys are checked here only statically)."""

from __future__ import annotations

import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_one_json_line():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3",
                          "--cpu-games", "8192", "--ref-kind", "port"], check=True, capture_output=True, text=True, cwd=ROOT).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 3 and d["warmup"] >= 3
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None


def test_reference_arm_times_the_staged_numba_reference():
    """`--impl reference` = the reference's own Numba VecGame (oracle/_ref) when it is staged and numba imports; its `config`
    is the GPU arm's config for the same command line (the driver compares them), the sample is described in cpu_baseline."""
    import pytest

    sys.path.insert(0, ROOT)
    import bench
    from oracle import make_ref

    if not make_ref.available():
        pytest.skip("oracle/_ref is not staged")
    pytest.importorskip("numba")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "3", "--warmup", "3",
                          "--cpu-games", "32768"], check=True, capture_output=True, text=True, cwd=ROOT).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    cb = d["cpu_baseline"]
    assert cb["kind"] == "reference" and cb["games"] == 32768 and cb["cores"] >= 1 and cb["value"] == d["value"] > 0
    assert "game_numba.py unmodified" in cb["sample"] and cb["step_only_value"] >= cb["value"]
    assert d["cpu_port"]["kind"] == "port" and d["cpu_port"]["value"] > 0
    ns = type("A", (), {"games_per_gpu": 1 << 24, "burn_in": bench.BURN_IN})()
    assert d["config"] == bench.make_config(ns, 1)  # what the GPU arm prints for the default command line
    assert d["steps"] == 3 and d["e2e"]["value"] == d["value"]


def test_other_ranks_of_the_reference_arm_do_nothing():
    env = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "3"],
                       check=True, capture_output=True, text=True, cwd=ROOT, env=env)
    assert r.stdout.strip() == ""


def test_gpu_arm_emits_the_contract_keys():
    src = open(os.path.join(ROOT, "bench.py")).read()
    for key in ('"metric"', '"value"', '"unit"', '"n_gpus"', '"steps"', '"warmup"', '"ms_per_step"', '"higher_is_better"',
                '"scaling"', '"vs_baseline"', '"dtype"', '"data"', '"config"', '"clocks"', '"e2e"', '"gpu_launches"', '"roofline"',
                '"cpu_baseline"', '"h2d_bytes_per_step"', '"d2h_bytes_per_step"', '"bound"', '"achieved"', '"peak"', '"frac"',
                '"traffic"', '"workload"', '"traffic_source"', '"shard_check"', '"collective_ms"', '"reduces_timed"',
                '"pcie_gbs"', '"pcie_ceiling_gbs"', '"cpu_port"'):
        assert key in src, key
