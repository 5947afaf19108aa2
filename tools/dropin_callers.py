"""The reference's OWN rollout loop around three environments, timed per runner step (what VERDICT r1 item 5 asks for):

    eval_perf.py shape      VecRunner + ReplayRecorder (both callbacks), 65 536 games, as eval_perf.py:66-102 builds them
    run_train3.py shape     VecRunner + RunnerStats + the trainer's seven host->buffer copies, 4096 games (run_train3.py:82-157)

once over the reference Numba VecGame, once over the CUDA VecGame through its NumPy (drop-in) surface.  The policy is the
reference's RandomPolicy on the CPU (the CNN checkpoint is absent; its cost is the same on both sides and is reported
separately as `policy_us`).  The unmodified reference modules come from oracle/_ref (oracle/make_ref.py).

    python tools/dropin_callers.py [--out gpurun_out/dropin_callers_r02.json]
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch


def timed_loop(make_env, shape, refpkg, steps, warm):
    VecRunner, RunnerStats = refpkg["runner"].VecRunner, refpkg["runner"].RunnerStats
    ReplayRecorder, REPLAY_SPEC = refpkg["replay"].ReplayRecorder, refpkg["replay"].REPLAY_SPEC
    m = shape["games"]
    env = make_env(m)
    env.reset(1)
    env_time = [0.0]

    def timed_method(fn):
        def wrapper(*args, **kw):
            t0 = time.perf_counter()
            out = fn(*args, **kw)
            env_time[0] += time.perf_counter() - t0
            return out

        return wrapper

    for name in ("prepare", "observations", "step"):  # instance attributes shadow the class methods: the callers are untouched
        setattr(env, name, timed_method(getattr(env, name)))
    policy_time = [0.0]
    base = refpkg["RandomPolicy"](seed=3)

    class TimedPolicy(refpkg["Policy"]):
        def sample_actions(self, state, valid_actions, *, generator=None):
            t0 = time.perf_counter()
            out = base.sample_actions(state, valid_actions, generator=generator)
            policy_time[0] += time.perf_counter() - t0
            return out

    runner = VecRunner(env, 16, sample_device="cpu")
    if shape["kind"] == "eval_perf":
        rec = ReplayRecorder(m, m)  # eval_perf.py:69
        runner.add_callback(VecRunner.EVENT_PREPARED, rec.on_prepared)
        runner.add_callback(VecRunner.EVENT_STEPPED, rec.on_stepped)
    else:
        stats = RunnerStats()
        runner.add_callback(VecRunner.EVENT_STEPPED, stats.on_stepped)
        buffers = {k: torch.zeros((1, steps + warm, m) + sh, dtype=dt) for k, (sh, dt) in REPLAY_SPEC.items()}
        si = [0]

        def on_stepped(game, result, actions, action_log_probs):  # run_train3.py:125-155
            def copy(name, src, dtype=None):
                buffers[name][0, si[0], ...].copy_(torch.from_numpy(src).to(dtype=dtype))

            copy("state", result["prev_state"], torch.int8)
            copy("valid_actions", result["prev_valid_actions"], torch.bool)
            copy("next_state", result["state"], torch.int8)
            copy("next_valid_actions", result["valid_actions"], torch.bool)
            copy("reward", result["reward"], torch.float32)
            copy("terminated", result["terminated"], torch.bool)
            copy("step", result["step"], torch.int32)
            buffers["action"][0, si[0], ...].copy_(actions.detach())
            si[0] += 1

        runner.add_callback(VecRunner.EVENT_STEPPED, on_stepped)
    policy = TimedPolicy()
    t0 = time.perf_counter()
    runner.step_once(policy)  # the all-games first prepare(): 65 536 recorder lookups of game._data[slot]["id"]
    first = time.perf_counter() - t0
    for _ in range(warm - 1):
        runner.step_once(policy)
    policy_time[0] = 0.0
    env_time[0] = 0.0
    t0 = time.perf_counter()
    for _ in range(steps):
        runner.step_once(policy)
    total = time.perf_counter() - t0
    return {"first_step_ms": first * 1e3, "runner_step_us": total / steps * 1e6, "policy_us": policy_time[0] / steps * 1e6,
            "env_calls_us": env_time[0] / steps * 1e6,  # prepare() + observations() + step() as the runner calls them
            "env_and_callbacks_us": (total - policy_time[0]) / steps * 1e6}


def main():
    p = argparse.ArgumentParser()
    p.add_argument("--out", default="gpurun_out/dropin_callers_r02.json")
    a = p.parse_args()
    from oracle import make_ref

    gn = make_ref.import_reference()
    import ml2048.replay as replay
    import ml2048.runner as runner
    from ml2048.policy import Policy
    from ml2048.policy.random import RandomPolicy
    import numba

    import ml2048_b200

    numba.set_num_threads(max(1, (os.cpu_count() or 2) // 2))  # eval_perf.py:48
    refpkg = {"runner": runner, "replay": replay, "Policy": Policy, "RandomPolicy": RandomPolicy}
    out = {"numba_threads": numba.get_num_threads(), "cpu_count": os.cpu_count(), "gpu": torch.cuda.get_device_name(0)}
    for shape in ({"kind": "eval_perf", "games": 65536, "steps": 40, "warm": 4}, {"kind": "run_train3", "games": 4096, "steps": 200, "warm": 10},
                  {"kind": "run_train3", "games": 2048, "steps": 200, "warm": 10}):
        row = {}
        for name, make in (("reference_numba", lambda m: gn.VecGame(m, gn.reward_fn_improved)),
                           ("ml2048_b200", lambda m: ml2048_b200.VecGame(m, gn.reward_fn_improved))):
            torch.manual_seed(0)
            row[name] = timed_loop(make, shape, refpkg, shape["steps"], shape["warm"])
        row["speedup_env_and_callbacks"] = row["reference_numba"]["env_and_callbacks_us"] / row["ml2048_b200"]["env_and_callbacks_us"]
        row["speedup_env_calls"] = row["reference_numba"]["env_calls_us"] / row["ml2048_b200"]["env_calls_us"]
        out[f'{shape["kind"]}_M{shape["games"]}'] = row
        print(shape, json.dumps(row), flush=True)
    os.makedirs(os.path.dirname(a.out) or ".", exist_ok=True)
    with open(a.out, "w") as fh:
        json.dump(out, fh, indent=1)


if __name__ == "__main__":
    main()
