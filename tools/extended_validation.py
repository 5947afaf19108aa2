"""One-off large differential run (not part of the regular suite): CUDA environment vs the CPU oracle in lock step at
M = 2^20 games for 300 runner steps per configuration, every state field compared bit for bit every 50 steps.

    python tools/extended_validation.py [--out gpurun_out/extended_validation.txt]
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import ml2048_b200
from oracle import oracle as orc

CONFIGS = [("normal", 0.8, 1), ("improved", 0.8, 2), ("rank", 0.5, 3), ("maxcell", 0.9, 4), ("improved", 1.0, 5), ("normal", 0.0, 6)]


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--games", type=int, default=1 << 20)
    p.add_argument("--steps", type=int, default=300)
    p.add_argument("--out", default="gpurun_out/extended_validation.txt")
    a = p.parse_args()
    orc.load_lib().orc_set_num_threads(len(os.sched_getaffinity(0)))
    lines = []
    m, n = a.games, a.steps
    acts = np.empty(m, np.int64)
    total = 0
    for reward, two_prob, seed in CONFIGS:
        t0 = time.perf_counter()
        ref = orc.OracleVecGame(m, reward, two_prob=two_prob)
        ref.reset(seed)
        env = ml2048_b200.VecGame(m, reward, two_prob=two_prob, output="torch", onehot="u8")
        env.reset(seed)
        compared = 0
        for t in range(n):
            (i0,) = ref.prepare()
            (i1,) = env.prepare()
            assert i1.numel() == i0.size, (reward, t)
            ref.random_valid_actions(seed * 100000 + t, acts)
            if t % 7 == 3:
                acts[:: 13] = (acts[:: 13] + 1) % 4  # some wrong directions on purpose: the invalid-move path
            ref.step(acts)
            env.step(torch.from_numpy(acts).cuda())
            if t % 50 == 49 or t == n - 1:
                d = ref._data
                for name, dev in (("board", env.observations()[0]), ("valid_actions", env.observations()[1]), ("step", env._step),
                                  ("id", env._id), ("terminated", env._terminated), ("invalid", env._invalid), ("merged", env._merged)):
                    assert np.array_equal(dev.cpu().numpy(), d[name]), (reward, name, t)
                for name, dev in (("score", env._score), ("reward", env._reward)):
                    assert np.array_equal(dev.cpu().numpy().view(np.uint32), d[name].view(np.uint32)), (reward, name, t)
                assert np.array_equal(i1.cpu().numpy(), i0), (reward, "reset indices", t)
                oh = env.observations_onehot()
                assert bool((oh.sum(dim=1) == 1).all())
                compared += 1
        assert env._game_count == ref._game_count
        total += m * n
        lines.append(f"reward={reward:8s} two_prob={two_prob} seed={seed}: {m} games x {n} steps, {compared} full comparisons, "
                     f"{ref._game_count} games started, all fields bit-identical ({time.perf_counter()-t0:.1f} s)")
        print(lines[-1], flush=True)
    lines.append(f"TOTAL {total/1e9:.2f} G env-steps compared, 0 mismatches; GPU {torch.cuda.get_device_name(0)}")
    print(lines[-1])
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    open(a.out, "w").write("\n".join(lines) + "\n")


if __name__ == "__main__":
    main()
