"""Compile oracle/vecgame_oracle.c -> oracle/liboracle.so (gcc, OpenMP).  Test infrastructure."""

from __future__ import annotations

import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "vecgame_oracle.c")
OUT = os.path.join(_HERE, "liboracle.so")


def build(force: bool = False) -> str:
    if not force and os.path.exists(OUT) and os.path.getmtime(OUT) >= os.path.getmtime(SRC):
        return OUT
    cmd = ["gcc", "-O3", "-march=x86-64-v2", "-fopenmp", "-fPIC", "-shared", "-fvisibility=hidden", "-Wall", "-Wextra", "-o", OUT, SRC]
    subprocess.run(cmd, check=True)
    return OUT


if __name__ == "__main__":
    print(build(force=True))
