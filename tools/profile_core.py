"""Small driver for ncu: the core-only (no one-hot) step on M games, replay or philox, random-valid or given actions.

    python tools/profile_core.py [--games N] [--rng replay|philox] [--actions random|given] [--onehot f32|bf16|u8] [--steps K]
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ml2048_b200

p = argparse.ArgumentParser()
p.add_argument("--games", type=int, default=1 << 24)
p.add_argument("--rng", default="replay")
p.add_argument("--actions", default="random")
p.add_argument("--onehot", default=None)
p.add_argument("--steps", type=int, default=20)
p.add_argument("--burn-in", type=int, default=64)
p.add_argument("--merged", action="store_true")
p.add_argument("--fused", action="store_true", help="step_random(auto_reset=True): the auto-reset fused into the step kernel")
a = p.parse_args()

env = ml2048_b200.VecGame(a.games, output="torch", rng_mode=a.rng, onehot=a.onehot, track_merged=a.merged, sync_free=True)
env.reset(0)
for _ in range(a.burn_in):
    env.prepare()
    env.step_random()
acts = None
if a.actions == "given":
    snap = env.state_dict()
    acts = torch.empty((a.steps, a.games), dtype=torch.uint8, device="cuda")
    for t in range(a.steps):
        env.prepare()
        env.step_random(return_actions=True)
        acts[t].copy_(env._actions_out)
    env.load_state_dict(snap)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(3 * a.steps)]
t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0.record()
if a.fused:
    env.step_random(auto_reset=True)  # allocates the bookkeeping buffers, publishes the first counts
    torch.cuda.synchronize()
    t0.record()
for t in range(a.steps):
    ev[3 * t].record()
    if not a.fused:
        env.prepare()
    ev[3 * t + 1].record()
    if a.fused:
        env.step_random(auto_reset=True)
    elif acts is None:
        env.step_random()
    else:
        env.step(acts[t])
    ev[3 * t + 2].record()
t1.record()
torch.cuda.synchronize()
prep = sum(ev[3 * t].elapsed_time(ev[3 * t + 1]) for t in range(a.steps)) / a.steps
step = sum(ev[3 * t + 1].elapsed_time(ev[3 * t + 2]) for t in range(a.steps)) / a.steps
tot = t0.elapsed_time(t1) / a.steps
print(f"games={a.games} rng={a.rng} actions={a.actions} onehot={a.onehot} merged={a.merged} fused_reset={a.fused}: prepare {prep*1e3:.1f} us, step {step*1e3:.1f} us, "
      f"total {tot*1e3:.1f} us/step -> {a.games/tot/1e6:.2f} G env-steps/s; step kernel alone {a.games/step/1e6:.2f} G/s "
      f"= {59*a.games/step/1e6/6543.1*100:.1f}% of HBM peak at 59 B/step")
