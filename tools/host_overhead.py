"""Wall-clock host cost of one eager prepare()+step (small M, nothing to wait for on the GPU)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ml2048_b200

for sched in (0, 256):
    env = ml2048_b200.VecGame(2048, output="torch", onehot="f32", track_merged=False, sync_free=True)
    env.reset(0)
    if sched:
        env.schedule_ahead(sched)
    acts = torch.zeros(2048, dtype=torch.uint8, device="cuda")
    for _ in range(200):
        env.prepare()
        env.step_random()
    torch.cuda.synchronize()
    n = 2000
    t0 = time.perf_counter()
    for _ in range(n):
        env.prepare()
    t1 = time.perf_counter()
    for _ in range(n):
        env.step_random()
    t2 = time.perf_counter()
    for _ in range(n):
        env.step(acts)
    t3 = time.perf_counter()
    torch.cuda.synchronize()
    print(f"schedule_ahead={sched}: prepare {(t1-t0)/n*1e6:.1f} us, step_random {(t2-t1)/n*1e6:.1f} us, step(actions) {(t3-t2)/n*1e6:.1f} us per call (host)")
