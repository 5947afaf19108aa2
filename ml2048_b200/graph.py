"""CUDA-graph replay of multi-step rollouts (see ``VecGame.schedule_ahead`` for the device-resident random schedule
that makes the environment's launches free of per-step host work)."""

from __future__ import annotations

from typing import Any, Optional

import torch

from .vecgame import VecGame


class GraphedRollout:
    """``steps`` runner steps (prepare + step each) captured once as a CUDA graph and replayed with one launch.

    The per-step host random numbers come from the device-resident schedule (``VecGame.schedule_ahead``), so a
    replay involves no host work besides the graph launch; the schedule window is refilled between replays when
    it runs low.  ``steps`` must be even (boards and masks are ping-pong buffers: an even number of steps ends in
    the buffers the graph started from).  Policy: uniform over valid actions, chosen in-kernel
    (``actions=None``), a caller-owned CUDA tensor of actions that the caller rewrites between replays
    (``actions=tensor``, read by every captured step), or a torch policy INSIDE the graph: ``logits_fn(env)`` is
    captured together with the environment kernels and must return the (M,4) float32 logits of the current
    observations; the step kernel samples from them (``step_from_logits``).  With ``buffers`` (a
    ``runner.RolloutBuffers``) step t of a replay writes its transition into row ``[use_index, t]`` and the sampled
    action's log-probability into ``action_log_prob[use_index, t]`` -- the whole rollout of run_train3.py's epoch
    (run_train3.py:175-183) becomes one graph launch."""

    def __init__(self, env: VecGame, steps: int, *, window: Optional[int] = None, actions: Optional[torch.Tensor] = None,
                 return_actions: bool = False, logits_fn: Any = None, buffers: Any = None, use_index: int = 0,
                 auto_reset: bool = False):
        if steps <= 0 or steps % 2:
            raise ValueError(f"steps={steps}: must be a positive even number")
        if auto_reset and (logits_fn is not None or actions is not None or buffers is not None):
            raise ValueError("auto_reset=True fuses prepare() into step_random(): in-kernel random policy only, no buffers")
        self.env, self.steps = env, int(steps)
        self.window = int(window) if window else self.steps * 16
        if self.window < self.steps:
            raise ValueError("window must cover at least one replay")
        env.configure(sync_free=True)
        if env._sched_len and env._sched_pos < env._sched_len:
            raise RuntimeError("the environment already has a pending device schedule")
        # min_steps = window reserves one table slot per step, so a window is never cut short by table refreshes and
        # always holds a whole number of replays
        self.window -= self.window % self.steps
        env.schedule_ahead(self.window, min_steps=self.window)
        self.graph = torch.cuda.CUDAGraph()
        side = torch.cuda.Stream(device=env.device)
        side.wait_stream(torch.cuda.current_stream(env.device))
        pos, cur = env._sched_pos, env._cur
        if buffers is not None and buffers.shape[1] < self.steps:
            raise ValueError("buffers hold fewer steps than one replay")
        with torch.cuda.stream(side):
            if logits_fn is not None:
                with torch.no_grad():
                    for _ in range(3):  # warm the policy's kernels / workspaces outside the capture
                        logits_fn(env)
                side.synchronize()
            env._skip_id_check = True  # the id-range accounting is done per replay() (no device reads while capturing)
            if auto_reset:
                env._ensure_autoreset()  # allocate outside the capture
            with torch.cuda.graph(self.graph, stream=side):
                for t in range(self.steps):
                    if auto_reset:
                        env.step_random(return_actions=return_actions, auto_reset=True)
                        continue
                    env.prepare()
                    record = buffers.row(use_index, t) if buffers is not None else None
                    if logits_fn is not None:
                        with torch.no_grad():
                            logits = logits_fn(env).to(torch.float32).contiguous()
                        lp = buffers["action_log_prob"][use_index, t] if buffers is not None else None
                        env.step_from_logits(logits, log_prob_out=lp, record=record)
                    elif actions is None:
                        env.step_random(return_actions=return_actions or record is not None, record=record)
                    else:
                        env.step(actions, record=record)
                env._set_record(None)
            env._skip_id_check = False
        torch.cuda.current_stream(env.device).wait_stream(side)
        # capturing executed nothing on the device: rewind the host mirrors
        env._sched_pos = pos
        assert env._cur == cur

    def replay(self, times: int = 1) -> None:
        env = self.env
        for _ in range(times):
            if env._sched_pos + self.steps > env._sched_len:
                if env._sched_pos != env._sched_len:
                    raise RuntimeError("the device schedule was consumed outside this GraphedRollout")
                env.schedule_ahead(self.window, min_steps=self.window)
            env._check_id_range(self.steps)  # ids are int32: raise before the device counter could wrap
            self.graph.replay()
            env._sched_pos += self.steps
            # the replay changed the device state behind the host's back: whole-record lookups (`env._data[slot]`) must not be
            # served from a host snapshot taken before it, nor observations() from a cached copy
            env._state_epoch += 1
            env._obs_cache = None
