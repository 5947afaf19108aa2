"""BASELINE configs[4]: eval_perf.py semantics on the GPU environment -- play `rounds` games to termination and
report the max-tile distribution of the games with id < rounds (eval_perf.py:80-115; "first N games" is NOT
"first N to terminate", README.md:112-118).  The trained checkpoint is absent from the reference mount
(.MISSING_LARGE_BLOBS), so the policy here is the random-valid policy (policy/random.py), chosen in-kernel;
with it this measures the ENVIRONMENT side of eval_perf only.

    python tools/eval_random.py [--rounds 65536] [--batch-size 65536] [--rng replay|philox] [--seed 0]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ml2048_b200


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--rounds", type=int, default=65536)
    p.add_argument("--batch-size", type=int, default=65536)
    p.add_argument("--rng", default="replay")
    p.add_argument("--seed", type=int, default=0)
    p.add_argument("--graph-steps", type=int, default=32)
    p.add_argument("--out", default=None)
    a = p.parse_args()
    rounds = a.rounds
    batch = min(rounds, a.batch_size)  # eval_perf.py:66
    env = ml2048_b200.VecGame(batch, output="torch", rng_mode=a.rng, track_merged=False, sync_free=True)
    env.reset(a.seed)
    env.enable_episode_log(rounds)
    roll = ml2048_b200.GraphedRollout(env, a.graph_steps, window=a.graph_steps * 16)
    log = env.episode_log()
    t0 = time.perf_counter()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    runner_steps = 0
    while True:
        roll.replay(4)
        runner_steps += 4 * a.graph_steps
        if bool((log["max_tile"] > 0).all()):  # every game with id < rounds has finished
            break
    ev1.record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    mt = log["max_tile"].cpu().long()
    steps = log["steps"].cpu().double()
    score = log["score"].cpu().double()
    rows = []
    for k in sorted(set(mt.tolist()), reverse=True):
        sel = mt == k
        rows.append({"tile": 2 ** k, "count": int(sel.sum()), "share": float(sel.sum()) / rounds,
                     "mean_steps": float(steps[sel].mean()), "mean_score": float(score[sel].mean())})
    res = {"rounds": rounds, "batch_size": batch, "rng": a.rng, "policy": "random-valid (in-kernel)",
           "runner_steps": runner_steps, "env_steps": runner_steps * batch, "device_ms": ev0.elapsed_time(ev1), "wall_s": wall,
           "env_steps_per_s": runner_steps * batch / (ev0.elapsed_time(ev1) * 1e-3), "total_games_started": env._game_count,
           "mean_steps": float(steps.mean()), "mean_score": float(score.mean()), "distribution": rows}
    for r in rows:  # the table eval_perf.py prints (eval_perf.py:106-115)
        print(f"{r['tile']:6d}: {r['share']:7.2%}  count={r['count']:6d}  mean steps={r['mean_steps']:8.1f}  mean score={r['mean_score']:9.1f}")
    print(json.dumps({k: v for k, v in res.items() if k != "distribution"}))
    if a.out:
        json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
