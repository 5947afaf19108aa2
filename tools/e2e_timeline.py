"""Where the time of the host-buffer step (VecGame.step(host actions, fetch=...), M = 2^24) goes: per-slice device timeline
(kernels done / copies done, CUDA events on the pipeline's own streams) next to the wall clock of prepare() and step().

    python tools/e2e_timeline.py [--games 16777216] [--onehot f32|none] [--steps 12]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ml2048_b200

p = argparse.ArgumentParser()
p.add_argument("--games", type=int, default=1 << 24)
p.add_argument("--onehot", default="f32")
p.add_argument("--steps", type=int, default=12)
p.add_argument("--burn-in", type=int, default=200)
p.add_argument("--keys", default="state,valid_actions,reward,terminated")
a = p.parse_args()

m = a.games
env = ml2048_b200.VecGame(m, output="torch", onehot=None if a.onehot == "none" else a.onehot)
env.reset(0)
for _ in range(a.burn_in):
    env.step_random(auto_reset=True)
acts = torch.empty((a.steps + 3, m), dtype=torch.uint8, pin_memory=True)
sd = env.state_dict()
for t in range(a.steps + 3):
    env.prepare()
    env.step_random(return_actions=True)
    acts[t].copy_(env._actions_out, non_blocking=True)
torch.cuda.synchronize()
env.load_state_dict(sd)
env.configure(output="numpy", sync_free=False)
keys = tuple(a.keys.split(","))
for t in range(3):
    env.prepare()
    env.step(acts[t], fetch=keys)
tp = ts = 0.0
rows = []
for t in range(3, 3 + a.steps):
    env._pipe_trace = trace = []
    t0 = time.perf_counter()
    env.prepare()
    t1 = time.perf_counter()
    env.step(acts[t], fetch=keys)
    t2 = time.perf_counter()
    tp += t1 - t0
    ts += t2 - t1
    torch.cuda.synchronize()
    start = trace[0][2]
    rows.append([(hi - lo, start.elapsed_time(k), start.elapsed_time(c)) for lo, hi, k, c in trace[1:]] + [(t2 - t1) * 1e3])
env._pipe_trace = None
print(f"onehot={a.onehot} games={m} keys={keys}: prepare {tp / a.steps * 1e3:.3f} ms, step {ts / a.steps * 1e3:.3f} ms (wall clock, mean of {a.steps})")
last = rows[-1]
print("last step: slice games, kernels done at (ms after the first launch), copies done at, GB/s of the slice's copies")
prev = 0.0
per_game = getattr(env, "last_step_d2h_bytes", 0) / m
for n, k, c in last[:-1]:
    print(f"  {n:9d}  {k:7.3f}  {c:7.3f}  {n * per_game / max(c - max(prev, k), 1e-6) / 1e6:6.1f}")
    prev = c
print(f"  step() returned after {last[-1]:.3f} ms; all copies done at {np.mean([r[-2][2] for r in rows]):.3f} ms (mean), step wall {np.mean([r[-1] for r in rows]):.3f} ms")
