"""BASELINE configs[1] end to end through the DROP-IN (NumPy) surface: what one runner step of the unmodified reference
costs around the environment -- prepare(), observations() -> host arrays, step(int64 host actions), and the seven result
fields Trainer.on_stepped reads (run_train3.py:138-149) -- for the CUDA environment and for the CPU port of the reference.

    python tools/dropin_latency.py [--games 4096] [--steps 400]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

FIELDS = ("prev_state", "prev_valid_actions", "state", "valid_actions", "reward", "terminated", "step")


def loop(env, steps, m):
    rng = np.random.default_rng(0)
    sink = 0
    t_total = 0.0
    for t in range(steps + 20):
        t0 = time.perf_counter()
        (idx,) = env.prepare()
        board, valid = env.observations()
        # stand-in policy on the host: first valid direction (argmax of the mask), like sample_actions().cpu().numpy()
        acts = np.asarray(valid).argmax(axis=1).astype(np.int64)
        res = env.step(acts)
        for k in FIELDS:
            sink += int(np.asarray(res[k]).reshape(-1)[0])
        if t >= 20:
            t_total += time.perf_counter() - t0
    return t_total / steps * 1e6, sink


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--games", type=int, default=4096)
    p.add_argument("--steps", type=int, default=400)
    a = p.parse_args()
    import ml2048_b200
    from oracle import oracle as orc

    for m in (a.games, 2048, 65536):
        gpu = ml2048_b200.VecGame(m, ml2048_b200.reward_fn_improved)
        gpu.reset(0)
        us_gpu, _ = loop(gpu, a.steps, m)
        cpu = orc.OracleVecGame(m, "improved")
        cpu.reset(0)
        orc.load_lib().orc_set_num_threads(len(os.sched_getaffinity(0)))
        us_cpu, _ = loop(cpu, a.steps, m)
        print(f"M={m}: drop-in runner step (prepare + observations + step + 7 result fields on the host): "
              f"CUDA env {us_gpu:.0f} us = {m/us_gpu:.1f} M env-steps/s | CPU port {us_cpu:.0f} us = {m/us_cpu:.1f} M env-steps/s")


if __name__ == "__main__":
    main()
