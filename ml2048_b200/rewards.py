"""Reward selectors.

The reference passes ``reward_fn`` as a Numba-jitted callable (reference: src/ml2048/game_numba.py:
408-504, used at :728).  An arbitrary jitted callback cannot cross into a CUDA kernel, so the four
functions the reference ships are selected BY IDENTITY/NAME and mapped to a kernel enum; anything else
raises ``ValueError``.  The reference's own function objects (``ml2048.game_numba.reward_fn_*``) are
accepted too, by name.
"""

from __future__ import annotations

from typing import Any

from . import _lib

_KINDS = {
    "normal": _lib.REWARD_NORMAL,
    "improved": _lib.REWARD_IMPROVED,
    "rank": _lib.REWARD_RANK,
    "maxcell": _lib.REWARD_MAXCELL,
}


class _RewardSelector:
    def __init__(self, name: str):
        self.__name__ = f"reward_fn_{name}"
        self.kind = _KINDS[name]

    def __call__(self, *args: Any, **kwargs: Any):
        raise TypeError(f"{self.__name__} is a selector for the CUDA kernel, not a host callable")

    def __repr__(self) -> str:
        return f"<ml2048_b200.{self.__name__}>"


reward_fn_normal = _RewardSelector("normal")      # game_numba.py:408-438
reward_fn_improved = _RewardSelector("improved")  # game_numba.py:441-466
reward_fn_rank = _RewardSelector("rank")          # game_numba.py:469-484
reward_fn_maxcell = _RewardSelector("maxcell")    # game_numba.py:487-504


def reward_kind(reward_fn: Any) -> int:
    if reward_fn is None:
        return _lib.REWARD_NORMAL  # game_numba.py:564-565
    if isinstance(reward_fn, _RewardSelector):
        return reward_fn.kind
    if isinstance(reward_fn, str):
        name = reward_fn
    else:
        name = getattr(reward_fn, "__name__", None) or getattr(getattr(reward_fn, "py_func", None), "__name__", "") or ""
    name = name.replace("reward_fn_", "")
    if name not in _KINDS:
        raise ValueError(
            f"reward_fn={reward_fn!r}: only reward_fn_normal/improved/rank/maxcell can run inside the CUDA kernel"
        )
    return _KINDS[name]
