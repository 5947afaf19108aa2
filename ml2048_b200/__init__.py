"""ml2048_b200 -- B200-native drop-in for ml2048's vectorised 2048 environment ``VecGame``.

Reference surface kept (reference: src/ml2048/game_numba.py): ``VecGame``, ``VecStepResult``,
``reward_fn_normal/improved/rank/maxcell`` and the action constants of src/ml2048/game.py:14-17.
"""

from .rewards import reward_fn_improved, reward_fn_maxcell, reward_fn_normal, reward_fn_rank
from .graph import GraphedRollout
from .vecgame import VecGame, VecStepResult

STEP_LEFT, STEP_RIGHT, STEP_UP, STEP_DOWN = 0, 1, 2, 3  # src/ml2048/game.py:14-17

__all__ = [
    "VecGame",
    "GraphedRollout",
    "VecStepResult",
    "reward_fn_normal",
    "reward_fn_improved",
    "reward_fn_rank",
    "reward_fn_maxcell",
    "STEP_LEFT",
    "STEP_RIGHT",
    "STEP_UP",
    "STEP_DOWN",
]
