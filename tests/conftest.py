"""pytest configuration: the `gpu` marker and shared fixture helpers."""

from __future__ import annotations

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")

# the library reads ML2048_PREPARE (fused / split auto-reset) once per process unless this is set before its first
# prepare(): the split-vs-fused test switches modes inside one process
os.environ.setdefault("ML2048_PREPARE_RECHECK", "1")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden(name: str) -> dict:
    """Load a fixture eagerly (NpzFile re-inflates an array on every [] access)."""
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


def golden_rollouts():
    return sorted(f for f in os.listdir(GOLDEN) if f.startswith("rollout_") and f.endswith(".npz"))


@pytest.fixture(scope="session")
def oracle():
    from oracle import oracle as orc

    orc.load_lib()
    return orc
