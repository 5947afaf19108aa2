"""Device-resident rollout driver: the GPU counterpart of the reference's ``VecRunner`` + ``Trainer.on_stepped``.

In the reference one runner step crosses the host/device boundary three times (reference:
src/ml2048/runner.py:74-109): observations are copied up for the policy (:91-95), the sampled actions
come back down for ``env.step`` (:104), and the ``stepped`` callback copies seven result arrays up again
into the ``(use, step, game)`` training buffers (run_train3.py:138-155).  Here the environment, the
observations, the actions and the training buffers all live on the GPU:

    env.prepare()  ->  policy(state, valid)  ->  env.step(actions, record=row)   # no host copies

and the transition row is written by the step kernel itself (``record=``).  The class keeps the reference
runner's surface -- ``DeviceRunner(env, capacity, sample_device=...)``, ``add_callback(event, fn)``,
``step_once(policy)``, ``step_many(policy, n)``, events ``"prepared"`` with ``(env, new_indices)`` and
``"stepped"`` with ``(env, result, actions, log_probs)`` -- so a ``Policy`` written for the reference
(``sample_actions(state: long (M,16), valid_actions: bool (M,4)) -> (actions, log_probs)``,
policy/__init__.py:10-43) works unchanged.
"""

from __future__ import annotations

from typing import Any, Callable, Optional

import torch

from .vecgame import VecGame, stats_to_dict

# Trajectory variables and their shapes (replay.py:10-20) + the advantage slot (run_train3.py ADV_SPEC)
REPLAY_SPEC = {
    "state": ((16,), torch.int8),
    "valid_actions": ((4,), torch.bool),
    "action": ((), torch.int8),
    "action_log_prob": ((), torch.float32),
    "reward": ((), torch.float32),
    "next_state": ((16,), torch.int8),
    "next_valid_actions": ((4,), torch.bool),
    "step": ((), torch.int32),
    "terminated": ((), torch.bool),
}
ADV_SPEC = {"adv": ((), torch.float32)}

_KERNEL_FIELDS = ("state", "valid_actions", "action", "reward", "next_state", "next_valid_actions", "step", "terminated")


class RolloutBuffers:
    """The ``(use, step, game)`` training buffers of ``Trainer.__init__`` (run_train3.py:112-123), on the device."""

    def __init__(self, use_count: int, step_count: int, game_count: int, device: Any, *, with_adv: bool = True):
        self.shape = (int(use_count), int(step_count), int(game_count))
        spec = dict(REPLAY_SPEC)
        if with_adv:
            spec.update(ADV_SPEC)
        self.tensors = {name: torch.zeros(self.shape + tail, dtype=dtype, device=device) for name, (tail, dtype) in spec.items()}

    def __getitem__(self, name: str) -> torch.Tensor:
        return self.tensors[name]

    def row(self, use_index: int, step_index: int) -> dict[str, torch.Tensor]:
        """The slices ``buffers[name][ui, si]`` the step kernel fills (contiguous: game is the fastest axis)."""
        return {name: self.tensors[name][use_index, step_index] for name in _KERNEL_FIELDS}

    def flat(self) -> dict[str, torch.Tensor]:
        """``(use*step*game, ...)`` views, as run_train3.py:220 feeds the learner."""
        n = self.shape[0] * self.shape[1] * self.shape[2]
        return {name: t.reshape((n,) + t.shape[3:]) for name, t in self.tensors.items()}


class DeviceRunner:
    """Run all games of a device-resident ``VecGame`` in lock step under a torch policy (runner.py:28-117)."""

    EVENT_PREPARED: str = "prepared"  # args: (env, new_indices)
    EVENT_STEPPED: str = "stepped"    # args: (env, result, actions, log_probs)

    def __init__(self, env: VecGame, capacity: int, *, sample_device: Any = None, buffers: Optional[RolloutBuffers] = None,
                 fused_sampler: bool = False):
        if env._output != "torch":
            raise ValueError("DeviceRunner needs VecGame(..., output='torch')")
        self.env = env
        self.sample_device = sample_device if sample_device is not None else env.device
        self._vec_size = env._size
        self._capacity = capacity
        self._listeners: dict[str, list[Callable[..., Any]]] = {self.EVENT_PREPARED: [], self.EVENT_STEPPED: []}
        self.buffers = buffers
        self.fused_sampler = bool(fused_sampler)
        self._use_index = 0
        self._step_index = 0
        self._log_prob = torch.zeros((self._vec_size,), dtype=torch.float32, device=env.device)

    def add_callback(self, event: str, fn: Callable[..., Any]) -> None:
        assert event in self._listeners, event  # runner.py:70
        self._listeners[event].append(fn)

    def _emit(self, event: str, *args: Any) -> None:
        for fn in self._listeners[event]:
            fn(*args)

    def set_slot(self, use_index: int, step_index: int = 0) -> None:
        """Where the next transition goes in ``buffers`` (the ``ui``/``si`` of run_train3.py:131-136)."""
        self._use_index, self._step_index = int(use_index), int(step_index)

    def step_once(self, policy: Any) -> None:
        env = self.env
        (new_indices,) = env.prepare()  # runner.py:78
        if self._listeners[self.EVENT_PREPARED]:
            self._emit(self.EVENT_PREPARED, env, new_indices)
        board, valid = env.observations()  # CUDA uint8 views: nothing to copy (runner.py:89-95)
        record = None
        log_prob_row = None
        if self.buffers is not None:
            if self._step_index >= self.buffers.shape[1]:
                raise RuntimeError("rollout buffers are full: call set_slot() for the next epoch")
            record = self.buffers.row(self._use_index, self._step_index)
            log_prob_row = self.buffers["action_log_prob"][self._use_index, self._step_index]
        with torch.no_grad():
            if self.fused_sampler:
                # the policy head's logits go straight into the step kernel, which samples, steps and records
                logits = policy.action_logits(board.to(self.sample_device, torch.long), valid.to(self.sample_device, torch.bool))
                logits = logits.to(env.device, torch.float32).contiguous()
                lp_out = log_prob_row if log_prob_row is not None else self._log_prob
                result = env.step_from_logits(logits, log_prob_out=lp_out, record=record)
                actions, log_probs = env.sampled_actions, lp_out
            else:
                actions, log_probs = policy.sample_actions(board.to(self.sample_device, torch.long),
                                                           valid.to(self.sample_device, torch.bool))  # runner.py:97-102
                result = env.step(actions.to(env.device), record=record)  # runner.py:104, no .cpu()
                if log_prob_row is not None:
                    log_prob_row.copy_(log_probs.detach())  # run_train3.py:152-155
        if self.buffers is not None:
            self._step_index += 1
        if self._listeners[self.EVENT_STEPPED]:
            self._emit(self.EVENT_STEPPED, env, result, actions, log_probs)

    def step_many(self, policy: Any, count: int) -> None:
        for _ in range(count):
            self.step_once(policy)


class DeviceRunnerStats:
    """``RunnerStats`` (runner.py:139-189) read from the counters the step kernel maintains: the max-tile
    histogram of finished games, no per-step host pass over the boards."""

    def __init__(self, env: VecGame):
        self.env = env

    def reset(self) -> None:
        self.env._stats_dev.zero_()

    @property
    def counts(self):
        return self.env.episode_stats()["max_tile_hist"]

    @property
    def terminated_count(self) -> int:
        return self.env.episode_stats()["episodes"]

    def summary(self) -> list[tuple]:
        counts = self.counts
        total = counts.sum()
        return [(2 ** power, int(counts[power]), counts[power] / total) for power in range(16, 0, -1) if counts[power]]

    def as_dict(self) -> dict:
        return stats_to_dict(self.env.episode_stats_tensor())


class UniformValidPolicy:
    """``RandomPolicy`` (policy/random.py:17-27) on the device: uniform over the valid actions."""

    def __init__(self, seed: int = 0):
        self._generator: Optional[torch.Generator] = None
        self._seed = seed

    def sample_actions(self, state: torch.Tensor, valid_actions: torch.Tensor, *, generator: Optional[torch.Generator] = None):
        if self._generator is None or self._generator.device != state.device:
            self._generator = torch.Generator(device=state.device)
            self._generator.manual_seed(self._seed)
        probs = valid_actions.float()
        none = probs.sum(dim=-1, keepdim=True) == 0
        probs = torch.where(none, torch.ones_like(probs), probs)
        actions = torch.multinomial(probs, 1, True, generator=generator or self._generator).squeeze(-1)
        log_probs = torch.log(probs.gather(-1, actions[:, None]).squeeze(-1) / probs.sum(dim=-1))
        return actions.long(), log_probs.float()

    def action_logits(self, state: torch.Tensor, valid_actions: torch.Tensor) -> torch.Tensor:
        return torch.zeros(valid_actions.shape, dtype=torch.float32, device=valid_actions.device)
