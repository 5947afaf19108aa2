"""Small, complete exercise of every kernel for compute-sanitizer (memcheck): odd sizes, all modes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ml2048_b200
from ml2048_b200.ops import encode_onehot, gae_advantages, sample_masked_categorical, valid_actions
from ml2048_b200.runner import RolloutBuffers

for m in (1, 17, 255, 4097, 70001):
    for rng in ("replay", "philox"):
        for onehot in (None, "f32", "bf16", "u8"):
            env = ml2048_b200.VecGame(m, ml2048_b200.reward_fn_maxcell, output="torch", rng_mode=rng, onehot=onehot, sync_free=True)
            env.reset(1)
            env.enable_episode_log(100)
            buf = RolloutBuffers(1, 4, m, "cuda")
            lp = torch.empty(m, device="cuda")
            for t in range(12):
                env.prepare()
                if t % 3 == 0:
                    env.step_random(return_actions=True, record=buf.row(0, t % 4))
                elif t % 3 == 1:
                    env.step_from_logits(torch.randn(m, 4, device="cuda"), log_prob_out=lp)
                else:
                    env.step(torch.randint(0, 4, (m,), device="cuda"))
            env.schedule_ahead(8)
            for t in range(10):
                env.prepare()
                env.step_random()
            env.summary()
            b = env.observations()[0]
            encode_onehot(b, torch.float32)
            valid_actions(b)
    logits = torch.randn(m, 4, device="cuda")
    sample_masked_categorical(logits, torch.rand(m, 4, device="cuda") < 0.5, seed=1, counter=1)
    v = torch.randn(2, 7, m, device="cuda")
    gae_advantages(v, v, v, v > 0, gamma=0.9, lambda_=0.9)
roll_env = ml2048_b200.VecGame(3000, output="torch", sync_free=True)
roll_env.reset(0)
ml2048_b200.GraphedRollout(roll_env, 4, window=16).replay(6)
host = ml2048_b200.VecGame(1 << 18, onehot="u8")
host.reset(0)
host.prepare()
host.step(np.zeros(1 << 18, np.uint8), fetch=("state", "reward"))
torch.cuda.synchronize()
print("sanitize smoke done")
