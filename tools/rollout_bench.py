"""Measurement of the rollout-path rows next to the environment at the run_train3.py shape
(4096 games x 16 steps x 2 buffers): runner step latency, transition recording, sampler and GAE against the torch
formulation the reference uses.

    python tools/rollout_bench.py [--out gpurun_out/rollout_bench.json]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ml2048_b200
from ml2048_b200.ops import gae_advantages, sample_masked_categorical
from ml2048_b200.runner import DeviceRunner, RolloutBuffers, UniformValidPolicy


def cuda_time(fn, n, warm=5):
    for _ in range(warm):
        fn()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3, (time.perf_counter() - t0) / n * 1e6  # device us, wall us


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--out", default="gpurun_out/rollout_bench.json")
    a = p.parse_args()
    res = {}
    m, steps = 4096, 16
    # --- runner step: env + policy + transition record -------------------------------------------------
    for name, fused in (("runner_step_policy_sample_actions", False), ("runner_step_fused_logits_sampler", True)):
        env = ml2048_b200.VecGame(m, ml2048_b200.reward_fn_improved, output="torch", sync_free=True, onehot="f32")
        env.reset(0)
        buf = RolloutBuffers(2, steps, m, "cuda")
        runner = DeviceRunner(env, steps, buffers=buf, fused_sampler=fused)
        policy = UniformValidPolicy(0)
        epoch = [0]

        def one_epoch():
            runner.set_slot(epoch[0] % 2, 0)
            runner.step_many(policy, steps)
            epoch[0] += 1

        dev, wall = cuda_time(one_epoch, 30)
        res[name] = {"device_us_per_step": dev / steps, "wall_us_per_step": wall / steps, "games": m,
                     "env_steps_per_s_wall": m / (wall / steps * 1e-6)}
    # --- the same transitions recorded the reference's way: 8 torch copies per step ----------------------
    env = ml2048_b200.VecGame(m, ml2048_b200.reward_fn_improved, output="torch", sync_free=True)
    env.reset(0)
    buf = RolloutBuffers(2, steps, m, "cuda")
    acts = torch.zeros(m, dtype=torch.uint8, device="cuda")
    si = [0]

    def step_with_torch_copies():
        env.prepare()
        r = env.step_random(return_actions=True)
        t = si[0] % steps
        buf["state"][0, t].copy_(r["prev_state"])
        buf["valid_actions"][0, t].copy_(r["prev_valid_actions"])
        buf["next_state"][0, t].copy_(r["state"])
        buf["next_valid_actions"][0, t].copy_(r["valid_actions"])
        buf["reward"][0, t].copy_(r["reward"])
        buf["terminated"][0, t].copy_(r["terminated"])
        buf["step"][0, t].copy_(r["step"])
        buf["action"][0, t].copy_(env.sampled_actions)
        si[0] += 1

    def step_with_kernel_record():
        env.prepare()
        t = si[0] % steps
        env.step_random(return_actions=True, record=buf.row(0, t))
        si[0] += 1

    for name, fn in (("record_by_8_torch_copies", step_with_torch_copies), ("record_by_step_kernel", step_with_kernel_record)):
        dev, wall = cuda_time(fn, 400, 20)
        res[name] = {"device_us_per_step": dev, "wall_us_per_step": wall}
    # --- sampler ----------------------------------------------------------------------------------------
    for mm in (4096, 1 << 22):
        logits = torch.randn(mm, 4, device="cuda")
        valid = torch.rand(mm, 4, device="cuda") < 0.7
        valid[:, 0] = True

        def ours():
            sample_masked_categorical(logits, valid, seed=1, counter=2)

        def torch_way():  # _sample_action, policy/actor_critic.py:56-76 + stats.py:32-47
            lg = torch.where(valid, logits, torch.finfo(torch.float32).min)
            dist = torch.distributions.Categorical(logits=lg)
            a_ = torch.multinomial(dist.probs, 1, True).squeeze(-1)
            dist.log_prob(a_)

        d0, w0 = cuda_time(ours, 100)
        d1, w1 = cuda_time(torch_way, 100)
        res[f"sampler_M{mm}"] = {"kernel_device_us": d0, "kernel_wall_us": w0, "torch_device_us": d1, "torch_wall_us": w1,
                                 "bytes_per_game": 16 + 4 + 8 + 4, "hbm_gbs": mm * 32 / (d0 * 1e-6) / 1e9}
    # --- GAE --------------------------------------------------------------------------------------------
    for (u, s, g) in ((2, 16, 4096), (2, 64, 1 << 20)):
        v0, v1 = torch.randn(u, s, g, device="cuda"), torch.randn(u, s, g, device="cuda")
        reward = torch.rand(u, s, g, device="cuda")
        term = torch.rand(u, s, g, device="cuda") < 0.01
        out = torch.empty_like(v0)

        def ours():
            gae_advantages(v0, v1, reward, term, gamma=0.997, lambda_=0.95, out=out)

        def torch_way():  # gae.py:50, :65-68
            mask = ~term
            delta = 0.997 * v1 * mask + reward - v0
            tmp = torch.zeros(u, g, device="cuda")
            for idx in reversed(range(s)):
                tmp = tmp * (0.997 * 0.95)
                tmp = delta[:, idx, :] + tmp * mask[:, idx, :]
                out[:, idx, :] = tmp

        d0, w0 = cuda_time(ours, 50)
        d1, w1 = cuda_time(torch_way, 20)
        n = u * s * g
        res[f"gae_{u}x{s}x{g}"] = {"kernel_device_us": d0, "kernel_wall_us": w0, "torch_device_us": d1, "torch_wall_us": w1,
                                   "bytes_per_transition": 17, "hbm_gbs": n * 17 / (d0 * 1e-6) / 1e9}
    for k, v in res.items():
        print(k, json.dumps(v))
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
