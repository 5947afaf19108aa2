"""BASELINE configs[1] and configs[2]: training-shape latency and the single-B200 throughput sweep.

    python tools/sweep.py [--out gpurun_out/sweep.json] [--quick]

Every point: reset(seed 0) -> 256-step warm-up (boards in steady state) -> 128 timed runner steps
(prepare + step with in-kernel random-valid actions), CUDA-event timed, once launching every kernel from
Python ("eager") and once replaying a 16-step CUDA graph fed by the device-resident schedule ("graph").
HBM fraction = algorithmic bytes per env step x steps/s / MEASURED_PEAKS.json hbm_gbs.
"""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

import ml2048_b200

BYTES = {None: 59, "f32": 59 + 1024, "bf16": 59 + 512, "u8": 59 + 256}


def hbm_peak():
    try:
        return float(json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")))["hbm_gbs"])
    except Exception:
        return 6650.0


def measure(m, rng, onehot, reward, timed=128, warm=256, graph_steps=16):
    out = {}
    for mode in ("eager", "graph", "fused_eager", "fused_graph"):  # fused_*: step_random(auto_reset=True), the reset inside the step kernel
        env = ml2048_b200.VecGame(m, reward, output="torch", rng_mode=rng, onehot=onehot, track_merged=False, sync_free=True)
        env.reset(0)
        for _ in range(warm):
            env.prepare()
            env.step_random()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if mode == "eager":
            for _ in range(8):
                env.prepare()
                env.step_random()
            torch.cuda.synchronize()
            a.record()
            for _ in range(timed):
                env.prepare()
                env.step_random()
            b.record()
        elif mode == "fused_eager":
            for _ in range(8):
                env.step_random(auto_reset=True)
            torch.cuda.synchronize()
            a.record()
            for _ in range(timed):
                env.step_random(auto_reset=True)
            b.record()
        else:
            roll = ml2048_b200.GraphedRollout(env, graph_steps, window=graph_steps * 16, auto_reset=mode == "fused_graph")
            roll.replay(1)
            torch.cuda.synchronize()
            a.record()
            roll.replay(timed // graph_steps)
            b.record()
        torch.cuda.synchronize()
        ms = a.elapsed_time(b) / timed
        out[mode] = {"us_per_step": ms * 1e3, "env_steps_per_s": m / (ms * 1e-3)}
        del env
        torch.cuda.empty_cache()
    return out


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--out", default="gpurun_out/sweep.json")
    p.add_argument("--quick", action="store_true")
    a = p.parse_args()
    peak = hbm_peak()
    res = {"hbm_peak_gbs": peak, "gpu": torch.cuda.get_device_name(0), "points": []}
    sizes = [1 << 16, 1 << 20, 1 << 24] if a.quick else [1 << 16, 1 << 18, 1 << 20, 1 << 22, 1 << 24]
    for m in sizes:
        for rng in ("philox", "replay"):
            for onehot in (None, "f32", "bf16", "u8"):
                if rng == "replay" and onehot in ("bf16", "u8"):
                    continue
                r = measure(m, rng, onehot, "normal")
                best = max(v["env_steps_per_s"] for v in r.values())
                pt = {"config": "sweep", "games": m, "rng": rng, "onehot": onehot, "bytes_per_step": BYTES[onehot], **r,
                      "hbm_frac_best": best * BYTES[onehot] / 1e9 / peak}
                res["points"].append(pt)
                print(json.dumps(pt), flush=True)
    # BASELINE configs[1]: the run_train3.py rollout shapes (BASELINE.json says 2048 x 64; the code has 4096 x 16)
    for m in (2048, 4096):
        r = measure(m, "replay", "f32", "improved", timed=128, warm=256, graph_steps=16)
        pt = {"config": "train_shape", "games": m, "rng": "replay", "onehot": "f32", "reward": "improved", **r}
        res["points"].append(pt)
        print(json.dumps(pt), flush=True)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    json.dump(res, open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
