"""Timing probe for the auto-reset at steady state: one prepare() per trial from the same snapshot, L2 flushed before.

    python tools/prepare_probe.py [--games N] [--onehot f32] [--trials 5]
ML2048_PREPARE=split|fused (read by the library at every call when ML2048_PREPARE_RECHECK is set) picks the three-launch or the single-launch path.
"""
import argparse
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ml2048_b200

p = argparse.ArgumentParser()
p.add_argument("--games", type=int, default=1 << 24)
p.add_argument("--onehot", default=None)
p.add_argument("--trials", type=int, default=5)
p.add_argument("--burn-in", type=int, default=256)
a = p.parse_args()
os.environ["ML2048_PREPARE_RECHECK"] = "1"  # the library otherwise reads the switch once per process
os.environ["ML2048_PREPARE"] = "split"  # the burn-in never runs the path under test
env = ml2048_b200.VecGame(a.games, output="torch", rng_mode="replay", onehot=a.onehot, track_merged=False, sync_free=True)
env.reset(0)
for _ in range(a.burn_in):
    env.prepare()
    env.step_random()
snap = env.state_dict()
flush = torch.empty(512 << 20, dtype=torch.uint8, device="cuda")
for mode in ("split", "fused", "split", "fused"):
    times = []
    for t in range(a.trials):
        env.load_state_dict(snap)
        flush.fill_(t)  # 512 MiB: nothing of the state is left in L2
        torch.cuda.synchronize()
        os.environ["ML2048_PREPARE"] = mode
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        env.prepare()
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1) * 1e3)
    print(f"{mode}: " + " ".join(f"{x:.1f}" for x in times) + " us")
