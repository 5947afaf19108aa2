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
                          "--cpu-games", "8192"], check=True, capture_output=True, text=True, cwd=ROOT).stdout
    lines = [l for l in out.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "env_steps_per_sec" and d["unit"] == "env-steps/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["steps"] == 3 and d["warmup"] >= 3
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and d["vs_baseline"] is None


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
                '"traffic"', '"workload"'):
        assert key in src, key
