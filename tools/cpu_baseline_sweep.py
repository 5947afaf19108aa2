"""BASELINE.md section 3 protocol for the CPU side, with the C port of the reference (the Numba reference itself cannot travel
to the GPU box): M in {1024, 2048, 4096, 2^16, 2^20, 2^22}, step()-only and prepare()+step(), all host threads and half.

    python tools/cpu_baseline_sweep.py [--out gpurun_out/cpu_baseline_sweep.json]
"""
import argparse
import json
import os
import platform
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from oracle import oracle as orc


def run(m, n, threads, lib):
    lib.orc_set_num_threads(threads)
    env = orc.OracleVecGame(m, "normal")
    env.reset(123)
    acts = np.empty(m, np.int64)
    env.prepare()
    env.random_valid_actions(0, acts)
    env.step(acts)
    best = {"step": 1e9, "both": 1e9}
    for rep in range(5):
        t_step = t_prep = 0.0
        for t in range(n):
            t0 = time.perf_counter()
            env.prepare()
            t1 = time.perf_counter()
            env.random_valid_actions(rep * 1000 + t, acts)
            t2 = time.perf_counter()
            env.step(acts)
            t3 = time.perf_counter()
            t_prep += t1 - t0
            t_step += t3 - t2
        best["step"] = min(best["step"], t_step)
        best["both"] = min(best["both"], t_step + t_prep)
    return {"games": m, "steps": n, "threads": threads, "step_only_Msteps_s": m * n / best["step"] / 1e6,
            "prepare_step_Msteps_s": m * n / best["both"] / 1e6}


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--out", default="gpurun_out/cpu_baseline_sweep.json")
    a = p.parse_args()
    lib = orc.load_lib()
    ncpu = len(os.sched_getaffinity(0))
    cpu = ""
    try:
        cpu = [l.split(":")[1].strip() for l in open("/proc/cpuinfo") if l.startswith("model name")][0]
    except Exception:
        cpu = platform.processor()
    res = {"cpu": cpu, "host_threads": ncpu, "impl": "oracle/vecgame_oracle.c (C port of the reference, OpenMP)", "points": []}
    for m, n in ((1024, 64), (2048, 64), (4096, 64), (1 << 16, 64), (1 << 20, 16), (1 << 22, 16)):
        for threads in (ncpu, max(1, ncpu // 2)):
            pt = run(m, n, threads, lib)
            res["points"].append(pt)
            print(json.dumps(pt), flush=True)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
