"""Where the host time of the NumPy drop-in surface goes at the training shape (M = 2048 / 4096): prepare() + step(numpy actions),
every result field read as the reference's callers do.  cProfile of the steady-state loop + wall clock per call.

    python tools/numpy_surface_profile.py [--games 2048] [--steps 3000] [--profile]
"""
import argparse
import cProfile
import os
import pstats
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

import ml2048_b200

p = argparse.ArgumentParser()
p.add_argument("--games", type=int, default=2048)
p.add_argument("--steps", type=int, default=3000)
p.add_argument("--profile", action="store_true")
a = p.parse_args()

env = ml2048_b200.VecGame(a.games)  # NumPy surface, reference defaults
env.reset(0)
rng = np.random.default_rng(0)
acts = rng.integers(0, 4, (64, a.games)).astype(np.int64)


def loop(n):
    tp = ts = 0.0
    for i in range(n):
        t0 = time.perf_counter()
        env.prepare()
        t1 = time.perf_counter()
        res = env.step(acts[i & 63])
        s = res["reward"].sum() + res["terminated"].sum() + res["state"][0, 0] + res["valid_actions"][0, 0]
        t2 = time.perf_counter()
        tp += t1 - t0
        ts += t2 - t1
    return tp / n * 1e6, ts / n * 1e6


loop(300)
tp, ts = loop(a.steps)
print(f"games={a.games}: prepare {tp:.1f} us, step+read {ts:.1f} us, total {tp + ts:.1f} us per runner step")
if a.profile:
    pr = cProfile.Profile()
    pr.enable()
    loop(a.steps)
    pr.disable()
    pstats.Stats(pr).sort_stats("tottime").print_stats(28)
