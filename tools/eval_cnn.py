"""BASELINE configs[4]: eval_perf.py with the CNN policy -- play `rounds` games to termination under
``CNNActorCriticPolicy(share_encoder=True)`` and report the max-tile distribution of the games with id < rounds
(eval_perf.py:60-115), with the environment, the observations, the policy and the sampling all on the GPU:

    env.prepare() -> fused fp32 one-hot (M,16,16) -> CNN actor -> logits -> step kernel samples, steps, logs the episode

The network below is CALLER CONTEXT, not product: a plain-torch restatement of the reference's actor path
(policy/_network.py:12-135 ``CNNEncoder``, :138-187 ``CNNActorNetwork``; module names match, so a reference checkpoint's
``policy_state`` loads with ``--save``) that starts from the fused one-hot instead of ``F.one_hot(x).float().permute``
(:86-95).  The trained checkpoint ``assets/ml2048_20240330_013340-epoch-2500.pt`` is absent from the reference mount
(.MISSING_LARGE_BLOBS); without ``--save`` the weights are the reference's own initialisation under
``torch.manual_seed(0)`` -- the same inference cost, an untrained policy's score distribution.

    python tools/eval_cnn.py [--rounds 65536] [--batch-size 65536] [--rng philox|replay] [--save ckpt.pt] [--out x.json]
"""
import argparse
import json
import math
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch import nn
from torch.nn import functional as F

import ml2048_b200

NUM_CLASSES = 16  # policy/_network.py:8


class CNNEncoder(nn.Module):
    """policy/_network.py:12-135, input = the class-major one-hot (N,16,16) the environment already produced."""

    def __init__(self, out_features: int = 1024, multiplier: int = 16):
        super().__init__()
        oc = out_features // 16
        self._out_channels = oc
        self._depthwise_full = nn.Conv1d(NUM_CLASSES, NUM_CLASSES * multiplier, NUM_CLASSES, groups=NUM_CLASSES)
        self._pointwise_full = nn.Conv1d(NUM_CLASSES * multiplier, oc * 4, 1)
        self._depthwise_hori = nn.Conv2d(NUM_CLASSES, NUM_CLASSES * multiplier, (1, 4), groups=NUM_CLASSES)
        self._pointwise_hori = nn.Conv2d(NUM_CLASSES * multiplier, oc, 1)
        self._depthwise_vert = nn.Conv2d(NUM_CLASSES, NUM_CLASSES * multiplier, (4, 1), groups=NUM_CLASSES)
        self._pointwise_vert = nn.Conv2d(NUM_CLASSES * multiplier, oc, 1)
        self._conv_out = nn.Conv1d(oc, out_features, 12)
        for m in (self._depthwise_full, self._depthwise_hori, self._depthwise_vert, self._conv_out):
            nn.init.zeros_(m.bias)  # :70-83

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        x_full = F.leaky_relu(self._pointwise_full(F.leaky_relu(self._depthwise_full(x))))
        board = x.reshape(-1, NUM_CLASSES, 4, 4)
        x_hori = F.leaky_relu(self._pointwise_hori(F.leaky_relu(self._depthwise_hori(board))))
        x_vert = F.leaky_relu(self._pointwise_vert(F.leaky_relu(self._depthwise_vert(board))))
        x = torch.cat((x_full.reshape(-1, self._out_channels, 4), x_hori.flatten(2), x_vert.flatten(2)), dim=2)
        return F.leaky_relu(self._conv_out(x)).flatten(1)


class CNNActorNetwork(nn.Module):
    """policy/_network.py:138-187."""

    def __init__(self, in_features: int = 1024, num_hidden: int = 256, num_hidden2: int = 64):
        super().__init__()
        self._fc1 = nn.Linear(in_features, num_hidden)
        self._fc2 = nn.Linear(num_hidden, num_hidden2)
        self._out = nn.Linear(num_hidden2, 4)
        for m, gain in ((self._fc1, math.sqrt(2)), (self._fc2, math.sqrt(2)), (self._out, 0.01)):
            nn.init.orthogonal_(m.weight, gain)
            nn.init.zeros_(m.bias)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        logits = self._out(F.relu(self._fc2(F.relu(self._fc1(x)))))
        return logits - logits.max(dim=-1, keepdim=True).values  # :181-185


class ActorPolicy(nn.Module):
    """The actor half of CNNActorCriticPolicy(share_encoder=True) (policy/actor_critic.py:240-284)."""

    def __init__(self):
        super().__init__()
        self._encoder = CNNEncoder(1024)
        self._actor = CNNActorNetwork(1024, 256, 64)

    def forward(self, onehot: torch.Tensor) -> torch.Tensor:
        return self._actor(self._encoder(onehot))


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--rounds", type=int, default=65536)
    p.add_argument("--batch-size", type=int, default=65536)
    p.add_argument("--rng", default="philox")
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--save", default=None, help="a reference checkpoint (torch.save dict with 'policy_state')")
    p.add_argument("--tf32", action="store_true", help="allow TF32 matmuls/convolutions (the reference runs plain fp32)")
    p.add_argument("--max-steps", type=int, default=200000)
    p.add_argument("--out", default=None)
    a = p.parse_args()
    torch.backends.cuda.matmul.allow_tf32 = a.tf32
    torch.backends.cudnn.allow_tf32 = a.tf32
    rounds = a.rounds
    batch = min(rounds, a.batch_size)  # eval_perf.py:66
    torch.manual_seed(0)
    policy = ActorPolicy().cuda().eval()
    weights = "reference initialisation, torch.manual_seed(0) (checkpoint absent)"
    if a.save:
        state = torch.load(a.save, map_location="cuda")["policy_state"]
        missing, unexpected = policy.load_state_dict(state, strict=False)
        assert not missing, missing  # the critic's tensors are `unexpected` here: the evaluation only needs the actor
        weights = os.path.basename(a.save)

    env = ml2048_b200.VecGame(batch, output="torch", rng_mode=a.rng, onehot="f32", track_merged=False, sync_free=True)
    env.reset(a.seed)
    env.enable_episode_log(rounds)
    log = env.episode_log()
    log_prob = torch.empty((batch,), dtype=torch.float32, device="cuda")

    def runner_step():
        env.prepare()
        with torch.no_grad():
            logits = policy(env.observations_onehot())
        env.step_from_logits(logits.contiguous(), log_prob_out=log_prob)

    for _ in range(3):  # warm-up: cuDNN algorithm selection, allocator
        runner_step()
    env.reset(a.seed)
    env._game_count = 0
    env.enable_episode_log(rounds)
    log = env.episode_log()
    torch.cuda.synchronize()

    ev = lambda: torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    start, stop = ev(), ev()
    start.record()
    probes = []  # (env-only, policy-only) CUDA-event pairs on a sample of the steps
    runner_steps = 0
    while runner_steps < a.max_steps:
        for _ in range(64):
            if runner_steps % 16 == 0:
                e0, e1, e2, e3 = ev(), ev(), ev(), ev()
                e0.record()
                env.prepare()
                e1.record()
                with torch.no_grad():
                    logits = policy(env.observations_onehot()).contiguous()
                e2.record()
                env.step_from_logits(logits, log_prob_out=log_prob)
                e3.record()
                probes.append((e0, e1, e2, e3))
            else:
                runner_step()
            runner_steps += 1
        if bool((log["max_tile"] > 0).all()):  # every game with id < rounds has finished
            break
    stop.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    dev_ms = start.elapsed_time(stop)
    env_us = sum(e0.elapsed_time(e1) + e2.elapsed_time(e3) for e0, e1, e2, e3 in probes) / len(probes) * 1e3
    pol_us = sum(e1.elapsed_time(e2) for e0, e1, e2, e3 in probes) / len(probes) * 1e3

    done = log["max_tile"] > 0
    mt = log["max_tile"].cpu().long()
    steps = log["steps"].cpu().double()
    score = log["score"].cpu().double()
    rows = []
    for k in sorted(set(mt[mt > 0].tolist()), reverse=True):
        sel = mt == k
        rows.append({"tile": 2 ** k, "count": int(sel.sum()), "share": float(sel.sum()) / rounds,
                     "mean_steps": float(steps[sel].mean()), "mean_score": float(score[sel].mean())})
    res = {"rounds": rounds, "batch_size": batch, "rng": a.rng, "policy": "CNNActorCriticPolicy(share_encoder=True) actor, fp32" + (" (TF32 allowed)" if a.tf32 else ""),
           "weights": weights, "finished": int(done.sum()), "runner_steps": runner_steps, "env_steps": runner_steps * batch,
           "device_ms": dev_ms, "wall_s": wall, "env_steps_per_s": runner_steps * batch / (dev_ms * 1e-3),
           "us_per_runner_step": dev_ms * 1e3 / runner_steps, "env_us_per_runner_step": env_us, "policy_us_per_runner_step": pol_us,
           "total_games_started": env._game_count, "mean_steps": float(steps[mt > 0].mean()), "mean_score": float(score[mt > 0].mean()),
           "distribution": rows}
    for r in rows:  # the table eval_perf.py prints (eval_perf.py:104-115)
        print(f"{r['tile']:6d}: {r['share']:7.2%}  count={r['count']:6d}  steps={r['mean_steps']:8.1f}  score={r['mean_score']:9.1f}")
    print(json.dumps({k: v for k, v in res.items() if k != "distribution"}))
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
