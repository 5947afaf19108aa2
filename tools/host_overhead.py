"""Wall-clock host cost of one eager prepare()+step (small M, nothing to wait for on the GPU)."""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ml2048_b200

for sched, rng in ((0, 'replay'), (256, 'replay'), (0, 'philox')):
    env = ml2048_b200.VecGame(2048, output="torch", onehot="f32", track_merged=False, sync_free=True, rng_mode=rng)
    env.reset(0)
    if sched:
        env.schedule_ahead(sched)
    acts = torch.zeros(2048, dtype=torch.uint8, device="cuda")
    for _ in range(200):
        env.prepare()
        env.step_random()
    torch.cuda.synchronize()
    n = 2000
    tp = ts = ta = 0.0
    for i in range(n):
        t0 = time.perf_counter()
        env.prepare()
        t1 = time.perf_counter()
        if i % 2:
            env.step_random()
            ts += time.perf_counter() - t1
        else:
            env.step(acts)
            ta += time.perf_counter() - t1
        tp += t1 - t0
    torch.cuda.synchronize()
    print(f"rng={rng} schedule_ahead={sched}: prepare {tp/n*1e6:.1f} us, step_random {ts/(n/2)*1e6:.1f} us, step(actions) {ta/(n/2)*1e6:.1f} us per call (host, refills included)")
