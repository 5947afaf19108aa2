"""The rollout half of run_train3.py's epoch (Trainer.loop_once, run_train3.py:175-218) on the device-resident path:
4096 games x 16 steps into the (use, step, game) buffers, then GAE -- with a small stand-in actor/critic (the
reference's CNN is out of scope; any torch module with the Policy surface works).  Prints per-stage times.

    python tools/train_shape_demo.py [--epochs 30]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch import nn

import ml2048_b200
from ml2048_b200.ops import gae_advantages
from ml2048_b200.runner import DeviceRunner, DeviceRunnerStats, RolloutBuffers


class TinyActorCritic(nn.Module):
    """Policy surface of the reference (policy/__init__.py:10-43): sample_actions / action_logits / eval_value."""

    def __init__(self):
        super().__init__()
        self.body = nn.Sequential(nn.Linear(256, 256), nn.LeakyReLU(), nn.Linear(256, 64), nn.LeakyReLU())
        self.actor = nn.Linear(64, 4)
        self.critic = nn.Linear(64, 1)

    def _features(self, state):
        x = ml2048_b200.ops.encode_onehot(state.to(torch.uint8).contiguous())  # (N,16,16) class-major, _network.py:86-95
        return self.body(x.flatten(1))

    def action_logits(self, state, valid_actions):
        return self.actor(self._features(state))

    def eval_value(self, state, valid_actions):
        return self.critic(self._features(state)).squeeze(-1)


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--epochs", type=int, default=30)
    a = p.parse_args()
    games, steps, uses = 4096, 16, 2  # run_train3.py:82-84
    env = ml2048_b200.VecGame(games, ml2048_b200.reward_fn_improved, output="torch", sync_free=True)
    env.reset(0)
    buf = RolloutBuffers(uses, steps, games, "cuda")
    runner = DeviceRunner(env, steps, buffers=buf, fused_sampler=True)
    policy = TinyActorCritic().cuda()
    stats = DeviceRunnerStats(env)
    t_roll = t_gae = 0.0
    for epoch in range(a.epochs + 3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        runner.set_slot(epoch % uses, 0)
        runner.step_many(policy, steps)                       # run_train3.py:175-183
        torch.cuda.synchronize()
        t1 = time.perf_counter()
        with torch.no_grad():                                  # gae.py:40-50
            flat = buf.tensors["state"].reshape(-1, 16)
            v0 = policy.eval_value(flat, None).reshape(uses, steps, games)
            v1 = policy.eval_value(buf.tensors["next_state"].reshape(-1, 16), None).reshape(uses, steps, games)
        gae_advantages(v0.contiguous(), v1.contiguous(), buf["reward"], buf["terminated"], gamma=0.997, lambda_=0.95, out=buf["adv"])
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        if epoch >= 3:
            t_roll += t1 - t0
            t_gae += t2 - t1
    n = a.epochs
    print(f"rollout: {t_roll/n*1e3:.2f} ms per epoch ({games*steps} transitions) = {t_roll/n/steps*1e6:.0f} us per runner step; "
          f"values+GAE: {t_gae/n*1e3:.2f} ms; finished episodes so far: {stats.terminated_count}")
    print("max-tile summary:", stats.summary()[:4])

    # the same rollout (environment + policy + sampling + recording) captured once as a CUDA graph
    env2 = ml2048_b200.VecGame(games, ml2048_b200.reward_fn_improved, output="torch", sync_free=True)
    env2.reset(0)
    buf2 = RolloutBuffers(1, steps, games, "cuda")

    def logits_fn(e):
        return policy.action_logits(e.observations()[0], None)

    roll = ml2048_b200.GraphedRollout(env2, steps, window=steps * 16, logits_fn=logits_fn, buffers=buf2)
    roll.replay(3)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    roll.replay(a.epochs)
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / a.epochs
    print(f"rollout as ONE CUDA graph: {dt*1e3:.2f} ms per epoch = {dt/steps*1e6:.0f} us per runner step")


if __name__ == "__main__":
    main()
