"""
oracle/make_ref.py -- stage the UNMODIFIED reference Python package under oracle/_ref/ (TEST INFRASTRUCTURE).

The reference (tsangwpx/ml2048) is a mounted Python tree with no setup.py/pyproject.toml, so it cannot be
pip-installed into baseline/_ref; and /root/reference does not exist on the GPU box.  This recipe copies the
few reference modules the CPU arm and the caller-parity tests import

    ml2048/__init__.py  game.py  game_numba.py          the Numba VecGame itself (the timed CPU baseline)
    ml2048/runner.py  replay.py  event.py  stats.py      its own callers (VecRunner, RunnerStats, ReplayRecorder)
    ml2048/policy/__init__.py  policy/random.py          Policy base class + RandomPolicy

byte for byte into oracle/_ref/ml2048/ during ``__graft_entry__.build()``.  oracle/_ref/ is git-ignored (the
sources stay out of this repository's history) but NOT gpurun-ignored, so it travels to the GPU box like the
built .so files.  A MANIFEST with the sha256 of every staged file is written next to them; the loader
refuses a tree whose hashes do not match it.

Only tests/, __graft_entry__ and bench.py (cpu_baseline / --impl reference) may import from oracle/_ref.
"""

from __future__ import annotations

import hashlib
import json
import os
import shutil
import sys
from typing import Any

_HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = "/root/reference/src/ml2048"
REF_DST = os.path.join(_HERE, "_ref")
PKG_DST = os.path.join(REF_DST, "ml2048")
MANIFEST = os.path.join(REF_DST, "MANIFEST.json")

FILES = (
    "__init__.py",
    "game.py",
    "game_numba.py",
    "runner.py",
    "replay.py",
    "event.py",
    "stats.py",
    "policy/__init__.py",
    "policy/random.py",
)


def _sha(path: str) -> str:
    with open(path, "rb") as fh:
        return hashlib.sha256(fh.read()).hexdigest()


def stage(force: bool = False) -> str | None:
    """Copy the reference modules into oracle/_ref/ml2048.  Returns the staged directory, or None when the
    reference mount is absent (the GPU box: the tree staged in the authoring container is used as is)."""
    if not os.path.isdir(REF_SRC):
        return REF_DST if available() else None
    if not force and available():
        try:
            with open(MANIFEST) as fh:
                have = json.load(fh)["files"]
            if all(have.get(f) == _sha(os.path.join(REF_SRC, f)) for f in FILES):
                return REF_DST
        except Exception:
            pass
    if os.path.isdir(PKG_DST):
        shutil.rmtree(PKG_DST)
    manifest = {}
    for rel in FILES:
        src, dst = os.path.join(REF_SRC, rel), os.path.join(PKG_DST, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        os.chmod(dst, 0o644)
        manifest[rel] = _sha(dst)
    with open(MANIFEST, "w") as fh:
        json.dump({"source": REF_SRC, "files": manifest}, fh, indent=1, sort_keys=True)
    return REF_DST


def available() -> bool:
    """True when a complete, unmodified copy is staged (hashes match the manifest written at staging time)."""
    try:
        with open(MANIFEST) as fh:
            have = json.load(fh)["files"]
        return all(_sha(os.path.join(PKG_DST, f)) == have[f] for f in FILES)
    except Exception:
        return False


def import_reference() -> Any:
    """Import the staged reference package and return its ``game_numba`` module.  Raises ImportError when the tree is
    not staged or numba is missing.  The staged tree shadows nothing: it is inserted at the FRONT of sys.path only if
    no ``ml2048`` package is imported yet."""
    if not available():
        raise ImportError("oracle/_ref is not staged: run `python -m oracle.make_ref` where /root/reference is mounted")
    os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
    sys.dont_write_bytecode = True
    mod = sys.modules.get("ml2048")
    if mod is not None and not os.path.abspath(getattr(mod, "__file__", "") or "").startswith(REF_DST):
        # already imported from somewhere else (a test that reads the live mount in the authoring container): accept it
        # only if it is byte-identical to the staged copy
        with open(MANIFEST) as fh:
            have = json.load(fh)["files"]
        other = os.path.join(os.path.dirname(os.path.abspath(mod.__file__)), "game_numba.py")
        if not os.path.exists(other) or _sha(other) != have["game_numba.py"]:
            raise ImportError(f"another `ml2048` package is already imported from {mod.__file__}")
    elif REF_DST not in sys.path:
        sys.path.insert(0, REF_DST)
    import ml2048.game_numba as game_numba  # noqa: PLC0415

    return game_numba


if __name__ == "__main__":
    print(stage(force=True))
